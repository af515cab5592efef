bash tools/gpu_final_tests.sh
bash tools/gpu_profile_round.sh
