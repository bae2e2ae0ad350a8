set -x
cd $GRAFT_REPO_ROOT
python bench.py > gpurun_out/r1_bench_1gpu.json 2> gpurun_out/r1_bench_1gpu.err
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r1_bench_reference.json 2>> gpurun_out/r1_bench_1gpu.err
timeout 300 python tools/bench_configs.py > gpurun_out/r1_configs_kernel_only.jsonl 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stream_kernel -s 4 -c 1 -o gpurun_out/r1_stream_kernel -f python tools/run_once.py config2 6 > /dev/null 2>&1
ls -la gpurun_out
