cd tests/native
for w in 1 2; do echo "waves=$w"; for cfg in "512 118 118 64 64 3 1" "512 59 59 128 128 3 1" "512 30 30 256 256 3 1" "512 118 118 64 256 1 1" "512 118 118 256 64 1 1" "512 15 15 512 512 3 1" "512 30 30 256 1024 1 1"; do T2R_WGRAD_WAVES=$w timeout 120 ./test_kernels bench $cfg 5 w | tail -1; done; done
