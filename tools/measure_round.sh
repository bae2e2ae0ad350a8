# One gpurun call that refreshes the round's measurements (B200 only): `sh tools/measure_round.sh r2`.
# Outputs land in gpurun_out/<round>_*; tools/ncu_summary.py turns the .ncu-rep files into the texts kept under profiles/.
set -x
cd $GRAFT_REPO_ROOT
r=${1:-r2}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${r}_gpu_tests.txt 2>&1
python bench.py > gpurun_out/${r}_bench_1gpu.json 2> gpurun_out/${r}_bench_1gpu.err
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/${r}_bench_reference_arm.json 2>> gpurun_out/${r}_bench_1gpu.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${r}_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-configs --no-sharded > gpurun_out/bench_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stream_kernel -s 4 -c 1 -o gpurun_out/${r}_stream_kernel -f python tools/run_once.py config2 6 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stream_kernel -s 4 -c 1 -o gpurun_out/${r}_stream_kernel_sparse -f python tools/run_once.py sparse 6 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"exact_window_kernel" -s 1 -c 1 -o gpurun_out/${r}_exact_window_kernel -f python tools/run_once.py config2 3 f64 > /dev/null 2>&1
GCRF_WIRE_SLICES=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:wire_decode -s 2 -c 1 -o gpurun_out/${r}_wire_decode_kernel -f python tools/e2e_probe.py > /dev/null 2>&1
timeout 200 python tools/e2e_probe.py > gpurun_out/${r}_e2e_probe.txt 2>&1
timeout 200 python tools/wire_slices_probe.py > gpurun_out/${r}_wire_slices.txt 2>&1
timeout 200 python tools/wire_slices_probe.py sparse >> gpurun_out/${r}_wire_slices.txt 2>&1
GCRF_WIRE_TRACE=1 timeout 200 python tools/wire_trace.py 2>&1 | tail -21 > gpurun_out/${r}_wire_trace.txt
timeout 200 python tools/segments_time.py > gpurun_out/${r}_segments_time.txt 2>&1
timeout 200 python tools/window_sizes_time.py > gpurun_out/${r}_window_sizes_time.txt 2>&1
timeout 200 python tools/density_probe.py > gpurun_out/${r}_density_probe.txt 2>&1
timeout 200 python tools/predict_tables_time.py 300000 3.0 > gpurun_out/${r}_predict_tables_time.txt 2>&1
timeout 200 python tools/dropin_time.py > gpurun_out/${r}_entry_levels_mibig.txt 2>&1
compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_f64.py tests/test_wire.py tests/test_refine.py -m gpu -x -q -k "golden or ragged_edge_cases_any_window or plain_call or segments_match or staging_area or being_timed or three_byte" > gpurun_out/${r}_sanitizer.txt 2>&1
ls -la gpurun_out | tail -30
