#!/bin/bash
# run bench.py against several builds of the library (tuning variants under build/): prints stage times; the JSON lines are
# kept in gpurun_out/sweep_<lib>.json
ARGS=${ARGS:---res 256 --spp 256 --steps 3 --warmup 1 --no-e2e --no-cpu-baseline --no-other-configs}
mkdir -p gpurun_out
for lib in "$@"; do
  tag=$(basename $lib .so)
  IA_B200_LIB=$PWD/$lib timeout -s KILL 300 python bench.py $ARGS 2>/dev/null | tail -1 > gpurun_out/sweep_$tag.json
  python -c "
import json,sys
d=json.loads(open('gpurun_out/sweep_$tag.json').read()); s=d['stages_ms']
print('$lib', 'value=%.3e'%d['value'], 'ms/step=%.1f'%d['ms_per_step'], 'shade=%.1f primary=%.1f occ=%.2f resample=%.2f setup=%.2f'%(s['shade'],s['primary'],s['occupancy'],s['resample'],s['setup']), 'spread', d.get('frame_cost_spread_ms'))"
done
