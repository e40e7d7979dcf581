bash tools/refresh_profiles.sh 2>&1 | tail -30
