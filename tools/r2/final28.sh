#!/bin/bash
# end-of-round validation of HEAD after the TIES rewrite: full GPU suite, smoke, default bench, reference arm, TIES traffic capture, sanitizer on TIES
mkdir -p gpurun_out
{
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -5
echo "=== smoke"
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "=== default bench"
SECONDS=0
timeout 1800 python bench.py > gpurun_out/r2_bench28.json 2> gpurun_out/r2_bench28.err
echo "rc=$? wall=${SECONDS}s"
tail -3 gpurun_out/r2_bench28.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench28.json').read().strip().splitlines()[-1])
def show(k,v):
    print(k, '|', v.get('metric'), v.get('value'), v.get('unit'), 'ms', v.get('ms_per_step'), 'frac', v.get('roofline',{}).get('frac'), 'traffic', v.get('roofline',{}).get('traffic'), 'e2e', v.get('e2e',{}).get('value') if v.get('e2e') else None)
show('primary', d)
for k in d:
    if isinstance(d[k], dict) and 'metric' in d[k]: show(k, d[k])
print('cpu_baseline', d.get('cpu_baseline'))
print('clocks', d.get('clocks'))
PY
echo "=== reference arm"
SECONDS=0
timeout 900 python bench.py --impl reference > gpurun_out/r2_bench28_ref.json 2> gpurun_out/r2_bench28_ref.err
echo "rc=$? wall=${SECONDS}s"; cut -c1-200 gpurun_out/r2_bench28_ref.json
echo "=== TIES traffic (one plan run = 6 launches; skip 3 warm-up runs)"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:ties_ -s 18 -c 6 --csv \
   --log-file gpurun_out/r2_ties28_traffic.csv python bench.py --workload ties > /dev/null 2>&1
grep -v "^==" gpurun_out/r2_ties28_traffic.csv | cut -d, -f5,13- | cut -c1-170 | tail -20
echo "=== compute-sanitizer on TIES (memcheck, racecheck)"
K="device_plan_bit_exact and bfloat16 and (3-gauss or 2-ints or 8-neg) or sampled_select or bracket_miss or fix_pass or beyond_the_fp16"
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_ties_gpu.py -x -q --timeout 800 -k "$K" 2>&1 | grep -v "Host Frame\|^=========$" | tail -6
timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_ties_gpu.py -x -q --timeout 800 -k "$K" 2>&1 | grep -v "Host Frame\|^=========$" | tail -8
} > gpurun_out/r2_final28.log 2>&1
tail -c 7000 gpurun_out/r2_final28.log
