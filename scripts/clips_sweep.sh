for c in 9 14 18; do
  timeout 600 python bench.py --clips $c --steps 4 --no-cpu-baseline --no-ref-cuda > gpurun_out/r3h_clips$c.json 2> gpurun_out/r3h_clips$c.err; echo "clips $c rc=$?"
  python -c "
import json
d=json.loads(open('gpurun_out/r3h_clips$c.json').read().strip().splitlines()[-1]); print('clips',$c,'value',round(d['value'],1),'ms',round(d['ms_per_step'],1),'e2e',round(d['e2e']['value'],1),'hbm GB',round(d['hbm_peak_bytes']/1e9,1))"
done
