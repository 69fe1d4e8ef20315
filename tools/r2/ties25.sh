#!/bin/bash
# TIES merge pass after the instruction diet (packed trim in 2 instructions, class-3 census by difference, paired 16-bit rounding,
# finalize folded into the fix-up kernel, init into the sampling kernel): parity, then timing, then a launch list.
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_ties_gpu.py -x -q --timeout 600 2>&1 | tail -5
echo "=== bench_ties"
for f in mean sum; do timeout 300 python tools/bench_ties.py --func $f 2>&1 | cut -c1-170; done
timeout 300 python tools/bench_ties.py --func max --kind neg 2>&1 | cut -c1-170
timeout 300 python tools/bench_ties.py --func mean --elements 320e6 2>&1 | cut -c1-170
timeout 300 python tools/bench_ties.py --func mean --dtype f16 2>&1 | cut -c1-170
timeout 300 python tools/bench_ties.py --func mean --src 8 --elements 80e6 2>&1 | cut -c1-170
echo "=== bench.py --workload ties"
timeout 600 python bench.py --workload ties 2>gpurun_out/r2_ties25.err | cut -c1-1200
tail -2 gpurun_out/r2_ties25.err
echo "=== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_ties25_launches.csv python tools/bench_ties.py --func mean --elements 320e6 --iters 2 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r2_ties25_launches.csv')) if len(r) > 10]
hdr = rows[0]; ki = hdr.index('Kernel Name'); mi = hdr.index('Metric Name'); vi = hdr.index('Metric Value')
t = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if r[mi] == 'gpu__time_duration.sum':
        t[r[ki][:70]][0] += 1; t[r[ki][:70]][1] += float(r[vi].replace(',', ''))
for k, (n, us) in sorted(t.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:72s} {n:4d} {us / n / 1e3 if us > 1e4 else us / n:10.1f}")
PY
} > gpurun_out/r2_ties25.log 2>&1
tail -c 6000 gpurun_out/r2_ties25.log
