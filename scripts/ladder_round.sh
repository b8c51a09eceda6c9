# BASELINE config 5 (ladder, block-Thomas engine): reductions with Cje = 1e-12 + exact Jacobian (convergence check), then the
# full size (256 stages, [16,16]) as a single ladder and as a batch of jittered ladders.  Every run under its own timeout.
TAG=${TAG:-lad}
timeout 100 python scripts/bench_ladder.py 256 2 2 1e-12 0.5 1 1 > gpurun_out/${TAG}_ladder256_2x2_cje.log 2>&1; tail -5 gpurun_out/${TAG}_ladder256_2x2_cje.log
timeout 100 python scripts/bench_ladder.py 32 4 4 1e-12 0.5 1 1 > gpurun_out/${TAG}_ladder32_4x4_cje.log 2>&1; tail -5 gpurun_out/${TAG}_ladder32_4x4_cje.log
for b in ${BATCHES:-4}; do
  timeout ${LADDER_TIMEOUT:-150} python scripts/bench_ladder.py 256 16 16 0 0.5 0 $b > gpurun_out/${TAG}_ladder_b$b.log 2>&1
  tail -5 gpurun_out/${TAG}_ladder_b$b.log
done
