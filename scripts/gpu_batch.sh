#!/bin/bash
# usage (on the GPU box, through gpurun): TAG=r2x [STEPS=...] bash scripts/gpu_batch.sh [tests] [bench] [fp16] [breakdown] [gma]
# Each stage writes gpurun_out/${TAG}_<stage>.{log,json}; stages are independent.
TAG=${TAG:-run}
mkdir -p gpurun_out
for stage in "$@"; do
  case $stage in
    tests) timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log ;;
    ktests) timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q > gpurun_out/${TAG}_pytest_kernels.log 2>&1; echo "ktests rc=$?"; tail -3 gpurun_out/${TAG}_pytest_kernels.log ;;
    bench) timeout 900 python bench.py ${BENCH_ARGS:-} > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err; echo "bench rc=$?"; python -c "
import json,sys
d=json.loads(open('gpurun_out/${TAG}_bench_default.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'parity',d.get('parity'),'issued_frac',d['roofline'].get('issued_frac'))" ;;
    quick) timeout 600 python bench.py --no-cpu-baseline --no-ref-cuda ${BENCH_ARGS:-} > gpurun_out/${TAG}_bench_quick.json 2> gpurun_out/${TAG}_bench_quick.err; echo "quick rc=$?"; tail -c 600 gpurun_out/${TAG}_bench_quick.json ;;
    fp16) timeout 600 python bench.py --precision fp16 --no-cpu-baseline --no-ref-cuda > gpurun_out/${TAG}_bench_fp16.json 2> gpurun_out/${TAG}_bench_fp16.err; echo "fp16 rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_bench_fp16.json').read().strip().splitlines()[-1]); print('fp16 value',d['value'],'parity',d.get('parity'))" ;;
    gma) timeout 900 python bench.py --ofe gma --clips 4 --no-cpu-baseline --no-ref-cuda > gpurun_out/${TAG}_bench_gma.json 2> gpurun_out/${TAG}_bench_gma.err; echo "gma rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_bench_gma.json').read().strip().splitlines()[-1]); print('gma value',d['value'],'parity',d.get('parity'))" ;;
    breakdown) CLIPS=${CLIPS:-9} timeout 600 python scripts/conv_breakdown.py > gpurun_out/${TAG}_conv_breakdown.txt 2>&1; echo "breakdown rc=$?"; head -40 gpurun_out/${TAG}_conv_breakdown.txt ;;
    ncu_conv) timeout 900 ncu --set full --import-source on --clock-control none --profile-from-start off -f -o gpurun_out/${TAG}_conv_targets python scripts/ncu_targets.py ${NCU_TARGETS:-gru_zr gru_q convc2 enc1 convc1} > gpurun_out/${TAG}_ncu_conv.log 2>&1; echo "ncu_conv rc=$?"; tail -3 gpurun_out/${TAG}_ncu_conv.log ;;
    ncu_aux) timeout 900 ncu --set full --import-source on --clock-control none --profile-from-start off -f -o gpurun_out/${TAG}_aux_targets python scripts/ncu_targets.py ${NCU_AUX:-lookup corr_gemm} > gpurun_out/${TAG}_ncu_aux.log 2>&1; echo "ncu_aux rc=$?"; tail -3 gpurun_out/${TAG}_ncu_aux.log ;;
    trace3) for x in 3 2 1; do TRACE_EXTRA=$x timeout 300 python scripts/mma_trace.py > gpurun_out/${TAG}_mma_trace_dbg$x.jsonl 2> gpurun_out/${TAG}_mma_trace_dbg$x.err; echo "== dbg $x rc=$?"; cat gpurun_out/${TAG}_mma_trace_dbg$x.jsonl; done ;;
    trace) timeout 300 python scripts/mma_trace.py > gpurun_out/${TAG}_mma_trace.jsonl 2> gpurun_out/${TAG}_mma_trace.err; echo "trace rc=$?"; cat gpurun_out/${TAG}_mma_trace.jsonl ;;
    probe) PROBE_PAIRS=${PROBE_PAIRS:-18} timeout 300 python scripts/gru_probe.py > gpurun_out/${TAG}_gru_probe.jsonl 2> gpurun_out/${TAG}_gru_probe.err; echo "probe rc=$?"; cat gpurun_out/${TAG}_gru_probe.jsonl ;;
    launches) timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python scripts/one_step.py > gpurun_out/${TAG}_launches.log 2>&1; echo "launches rc=$?"; python scripts/agg_launches.py gpurun_out/${TAG}_launches.csv 24 | tee gpurun_out/${TAG}_launches_summary.txt ;;
    lookup_ab) (python scripts/lookup_ab.py; ACCFLOW_LOOKUP=fast python scripts/lookup_ab.py) > gpurun_out/${TAG}_lookup_ab.jsonl 2> gpurun_out/${TAG}_lookup_ab.err; echo "lookup_ab rc=$?"; cat gpurun_out/${TAG}_lookup_ab.jsonl ;;
    narrow) for cfg in "128 592" "64 256" "64 592" "96 256"; do set -- $cfg; echo "== BN_CAP=$1 MSUB_MIN=$2"; ACCFLOW_TC_BN_CAP=$1 ACCFLOW_TC_MSUB_MIN=$2 PROBE_PAIRS=18 timeout 300 python scripts/gru_probe.py 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['conv'][:44].ljust(44), d['us_dbg0'], d['us_dbg32'])
"; done > gpurun_out/${TAG}_narrow.txt 2>&1; cat gpurun_out/${TAG}_narrow.txt ;;
    bf16) timeout 600 python bench.py --precision bf16 --no-ref-cuda > gpurun_out/${TAG}_bench_bf16.json 2> gpurun_out/${TAG}_bench_bf16.err; echo "bf16 rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_bench_bf16.json').read().strip().splitlines()[-1]); print('bf16 value',d['value'],'e2e',d['e2e']['value'],'parity',d.get('parity'))" ;;
    fp16full) timeout 600 python bench.py --precision fp16 --no-ref-cuda > gpurun_out/${TAG}_bench_fp16.json 2> gpurun_out/${TAG}_bench_fp16.err; echo "fp16full rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_bench_fp16.json').read().strip().splitlines()[-1]); print('fp16 value',d['value'],'e2e',d['e2e']['value'],'parity',d.get('parity'))" ;;
    gmafull) timeout 900 python bench.py --ofe gma --clips 4 --no-ref-cuda > gpurun_out/${TAG}_bench_gma.json 2> gpurun_out/${TAG}_bench_gma.err; echo "gmafull rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_bench_gma.json').read().strip().splitlines()[-1]); print('gma value',d['value'],'e2e',d['e2e']['value'],'parity',d.get('parity'))" ;;
    warm) timeout 600 python bench.py --warm-start --no-ref-cuda > gpurun_out/${TAG}_bench_warm_start.json 2> gpurun_out/${TAG}_bench_warm_start.err; echo "warm rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_bench_warm_start.json').read().strip().splitlines()[-1]); print('warm value',d['value'],'parity',d.get('parity'))" ;;
    memcheck) timeout 1200 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_kernels.py -q -x -m gpu > gpurun_out/${TAG}_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/${TAG}_sanitizer_memcheck.log ;;
    *) echo "unknown stage $stage" ;;
  esac
done
