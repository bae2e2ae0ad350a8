# One gpurun call that refreshes the round's measurements (B200 only).  "quick" skips what does not change when the
# streaming marginal kernel is untouched (reference arm, the other configs, the full ncu capture of that kernel).
set -x
cd $GRAFT_REPO_ROOT
mode=${1:-full}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1_gpu_tests.txt 2>&1
python bench.py > gpurun_out/r1_bench_1gpu.json 2> gpurun_out/r1_bench_1gpu.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
timeout 120 python tools/features_time.py > gpurun_out/r1_features_time.txt 2>&1
timeout 200 python tools/predict_tables_time.py 300000 3.0 > gpurun_out/r1_predict_tables_time.txt 2>&1
if [ "$mode" = full ]; then
  python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r1_bench_reference.json 2>> gpurun_out/r1_bench_1gpu.err
  timeout 300 python tools/bench_configs.py > gpurun_out/r1_configs_kernel_only.jsonl 2>&1
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:stream_kernel -s 4 -c 1 -o gpurun_out/r1_stream_kernel -f python tools/run_once.py config2 6 > /dev/null 2>&1
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:features_kernel -s 2 -c 1 -o gpurun_out/r1_features_kernel -f python tools/features_time.py > /dev/null 2>&1
fi
ls -la gpurun_out
