# ncu captures of the dense-block kernels of the staged engines (panel Gauss-Jordan, DMMA trailing update, DMMA Schur GEMM)
# on a 5-stage [16,16] ladder (S = 1089: the block size of BASELINE config 5); the same command first runs without ncu.
TAG=${TAG:-r02}
timeout 120 python scripts/bench_ladder.py 5 16 16 0 0.5 0 1 > gpurun_out/dense_plain_$TAG.log 2>&1 || exit 1
for k in k_gjp_update k_gjp_panel k_btd_schur; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$k -s 40 -c 1 -f -o gpurun_out/${k}_$TAG python scripts/bench_ladder.py 5 16 16 0 0.5 0 1 > gpurun_out/ncu_${k}_$TAG.log 2>&1
done
tail -4 gpurun_out/dense_plain_$TAG.log
