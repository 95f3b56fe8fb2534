# usage: bash scripts/conv_sweep.sh [kinds]   (runs from the repo root on the GPU box)
cd tests/native
K=${1:-fdw}
for cfg in "512 118 118 64 64 3 1" "512 59 59 128 128 3 1" "512 30 30 256 256 3 1" "512 15 15 512 512 3 1" "512 118 118 64 256 1 1" "512 118 118 256 64 1 1" "512 59 59 512 128 1 1" "512 30 30 256 1024 1 1" "512 30 30 1024 256 1 1" "512 15 15 2048 512 1 1" "512 79 79 64 64 5 1" "512 118 118 128 128 3 2"; do timeout 120 ./test_kernels bench $cfg 4 $K | tail -3; done
