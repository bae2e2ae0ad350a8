# Copies one round's measurements from gpurun_out/ (scratch) into profiles/ (tracked) and summarises the .ncu-rep files:
#   sh tools/collect_profiles.sh r2
r=${1:-r2}
cd "$(dirname "$0")/.."
for f in bench_1gpu.json bench_2gpu.json bench_8gpu.json bench_reference_arm.json launches.csv gpu_tests.txt sanitizer.txt \
         e2e_probe.txt wire_slices.txt wire_trace.txt segments_time.txt window_sizes_time.txt density_probe.txt predict_tables_time.txt entry_levels_mibig.txt topo_8gpu.txt; do
    [ -s gpurun_out/${r}_$f ] && cp gpurun_out/${r}_$f profiles/${r}_$f
done
for k in stream_kernel stream_kernel_sparse exact_window_kernel wire_decode_kernel; do
    [ -s gpurun_out/${r}_$k.ncu-rep ] && python tools/ncu_summary.py gpurun_out/${r}_$k.ncu-rep > profiles/${r}_ncu_summary_$k.txt
done
python - "$r" <<'PY'
import json, re, sys
r = sys.argv[1]
txt = open(f"profiles/{r}_ncu_summary_stream_kernel.txt").read()
val = lambda name: float(re.search(re.escape(name) + r" \[(\w*)\] = ([\d.,]+)", txt).group(2).replace(",", ""))
unit = lambda name: re.search(re.escape(name) + r" \[(\w*)\]", txt).group(1)
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
total = sum(val(n) * scale[unit(n)] for n in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
kernel = re.search(r"# kernel: (.*)", txt).group(1)
json.dump({"kernel": kernel, "genes": 2000810, "dram_bytes_per_launch": total,
           "source": f"profiles/{r}_ncu_summary_stream_kernel.txt (dram__bytes_read.sum + dram__bytes_write.sum of one launch)"},
          open("profiles/traffic_bytes_per_launch.json", "w"), indent=1)
print(open("profiles/traffic_bytes_per_launch.json").read())
PY
