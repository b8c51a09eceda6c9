# A/B of library builds on the bench workload (one gpurun call): TAG names the outputs, LIBS the variants (default + each)
TAG=${TAG:-ab}
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; tail -3 gpurun_out/${TAG}_tests.log
python scripts/ab_value.py >> gpurun_out/${TAG}_ab.log 2>&1
for v in $LIBS; do
  YHB_LIB=yalrf_b200/libyhb_$v.so python scripts/ab_value.py >> gpurun_out/${TAG}_ab.log 2>&1
done
for e in $ENVS; do
  env ${e//,/ } python scripts/ab_value.py >> gpurun_out/${TAG}_ab.log 2>&1
done
[ -f yalrf_b200/libyhb_timing.so ] && YHB_LIB=yalrf_b200/libyhb_timing.so python scripts/fused_timing.py > gpurun_out/${TAG}_timing.log 2>&1
grep -v "^yhb:" gpurun_out/${TAG}_ab.log | cut -c1-420
