# ncu --set full captures of the dominant kernels (one launch each), via the native bench modes.
cd tests/native
M="--set full --clock-control none --import-source on --launch-skip 1 --launch-count 1 -f"
ncu $M -k regex:conv_igemm_kernel -o ../../gpurun_out/r01g_igemm256_3x3_256 ./test_kernels bench 512 30 30 256 256 3 1 2 f > /dev/null 2>&1
ncu $M -k regex:conv_igemm_tma_kernel -o ../../gpurun_out/r01g_igemm_tma128_1x1_64_256_res ./test_kernels bench 512 118 118 64 256 1 1 2 f 2 > /dev/null 2>&1
ncu $M -k regex:conv_wgrad_kernel -o ../../gpurun_out/r01g_wgrad256_3x3_256 ./test_kernels bench 512 30 30 256 256 3 1 2 w > /dev/null 2>&1
ncu $M -k regex:conv_halo_kernel -o ../../gpurun_out/r01g_halo_3x3_64 ./test_kernels bench 512 118 118 64 64 3 1 2 f > /dev/null 2>&1
ncu $M -k regex:conv_wgrad_halo_kernel -o ../../gpurun_out/r01g_wgrad_halo_3x3_64 ./test_kernels bench 512 118 118 64 64 3 1 2 w > /dev/null 2>&1
ncu $M -k regex:conv_igemm_tma_kernel -o ../../gpurun_out/r01g_igemm_tma128_3x3_128 ./test_kernels bench 512 59 59 128 128 3 1 2 f > /dev/null 2>&1
ncu $M -k regex:bn_bwd_apply -o ../../gpurun_out/r01g_bn_bwd_apply_c256 ./test_kernels benchbn 7129088 256 1 > /dev/null 2>&1
ncu $M -k regex:bn_apply_rows -o ../../gpurun_out/r01g_bn_apply_c256 ./test_kernels benchbn 7129088 256 1 > /dev/null 2>&1
for f in ../../gpurun_out/r01g_*.ncu-rep; do
  echo "== $f"
  ncu -i $f --page raw --csv 2>/dev/null | python3 -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; v=rows[2] if len(rows)>2 else rows[1]
d=dict(zip(h,v))
keys=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sector_hit_rate.pct','dram__throughput.avg.pct_of_peak_sustained_elapsed','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active']
for k in keys:
    for kk in d:
        if kk.endswith(k) or kk==k: print(kk, d[kk]); break
"
done
