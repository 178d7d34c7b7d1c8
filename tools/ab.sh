#!/bin/bash
# usage: tools/ab.sh "<label> <env assignments>" ...   runs bench C5@16spp per variant and prints value + stage times
for v in "$@"; do
  label=${v%% *}; envs=${v#* }; [ "$envs" = "$v" ] && envs=""
  out=$(env $envs timeout 300 python bench.py --workload ${AB_WORKLOAD:-C5} --spp ${AB_SPP:-16} --steps 2 --warmup 1 --no-cpu-baseline 2>/dev/null | tail -1)
  echo "$out" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms_per_step_rank0']
print('$label', 'Mrays/s=%.1f'%d['value'], 'trace=%.1f'%s['trace_ms'], 'shade=%.1f'%s['shade_ms'], 'film=%.1f'%s['film_ms'], 'gen=%.1f'%s['generate_ms'], 'ms/step=%.1f'%d['ms_per_step'])
"
done
