# One GPU call that produces everything profiles/ quotes for a round (run through gpurun; TAG names the outputs):
#   gpurun --timeout 1500 -- 'TAG=r02a bash scripts/profile_round.sh'      then: python scripts/make_profiles.py r02a 'build'
# Every ncu pass runs after the same command has exited 0 without ncu; numbers printed under ncu are never bench values.
TAG=${TAG:-r02a}
set -x
python bench.py > gpurun_out/bench_$TAG.log 2> gpurun_out/bench_$TAG.err || exit 1
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref_$TAG.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_launch_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fused -c 1 -f -o gpurun_out/fused_$TAG python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_fused_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_dc_warp -c 1 -f -o gpurun_out/dc_$TAG python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_dc_$TAG.log 2>&1
python scripts/bench_device_eval.py > gpurun_out/deveval_$TAG.json 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_stage_device_eval -c 1 -f -o gpurun_out/deveval_$TAG python scripts/bench_device_eval.py > gpurun_out/ncu_deveval_$TAG.log 2>&1
python scripts/bench_configs.py > gpurun_out/configs_$TAG.json 2> gpurun_out/configs_$TAG.err
tail -c 600 gpurun_out/bench_$TAG.log; tail -3 gpurun_out/configs_$TAG.err
