cd tests/native
for mode in new old; do echo $mode; for cfg in "512 118 118 64 256 1 1 4 f 2" "512 59 59 128 512 1 1 4 f 2" "512 30 30 256 1024 1 1 4 f 2" "512 118 118 256 64 1 1 4 d 2" "512 118 118 256 128 1 1 4 d 2"; do if [ $mode = old ]; then T2R_DISABLE_TMA_EPI=1 timeout 120 ./test_kernels bench $cfg | tail -1; else timeout 120 ./test_kernels bench $cfg | tail -1; fi; done; done
