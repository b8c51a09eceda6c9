set -x
python bench.py > gpurun_out/bench_r01d.log 2> gpurun_out/bench_r01d.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r01d.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01d.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_launch_r01d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fused -c 1 -f -o gpurun_out/fused_r01d python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_fused_r01d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_dc_warp -c 1 -f -o gpurun_out/dc_r01d python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_dc_r01d.log 2>&1
python scripts/bench_configs.py > gpurun_out/configs_r01d.json 2> gpurun_out/configs_r01d.err
python scripts/bench_device_eval.py > gpurun_out/deveval_r01d.json 2>&1
tail -c 600 gpurun_out/bench_r01d.log; cat gpurun_out/deveval_r01d.json; tail -3 gpurun_out/configs_r01d.err
