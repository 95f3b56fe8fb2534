cd tests/native
./test_kernels bn 2>&1 | grep -v PASS | tail -3
for cfg in "7129088 256" "1782272 512" "460800 1024" "7129088 64"; do timeout 120 ./test_kernels benchbn $cfg 4 | tail -3; done
