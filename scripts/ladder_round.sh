# BASELINE config 5 (ladder, block-Thomas engine): parity tests of the engine, then the full size (256 stages, [16,16],
# Cje = 0: see DESIGN section 5) as a single ladder and as batches of jittered ladders.  Every run under its own timeout.
TAG=${TAG:-lad}
python -m pytest tests -m gpu -x -q -k "ladder or gjp or engines_agree" > gpurun_out/${TAG}_tests.log 2>&1; tail -3 gpurun_out/${TAG}_tests.log
for b in ${BATCHES:-1 4}; do
  timeout ${LADDER_TIMEOUT:-150} python scripts/bench_ladder.py 256 16 16 0 0.5 0 $b > gpurun_out/${TAG}_ladder_b$b.log 2>&1
  tail -5 gpurun_out/${TAG}_ladder_b$b.log
done
