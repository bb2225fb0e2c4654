# usage: bash tools/light_sweep.sh N EDITS "ENV1=.. ENV2=.." "ENV..." ...   (one light_bench run per environment string)
N=$1; E=$2; shift 2
for envs in "$@"; do
  env $envs timeout 300 python tools/light_bench.py $N $E 2>&1 | tail -1 > /tmp/lb.json
  python - "$envs" $N <<'PY'
import json, sys
d = json.loads(open('/tmp/lb.json').read())
i, a = d['initial_convergence'], d['after_edits']
print('%-60s %s^3 init %.2f M/s (%d upd, %.2fs)  edits %.2f M/s (%d upd, %.2fs, maxdiff %d)' % (sys.argv[1], sys.argv[2], i['cube_updates_per_s'] / 1e6, i['cube_updates'], i['seconds'], a['cube_updates_per_s'] / 1e6, a['cube_updates'], a['seconds'], a['max_difference']))
PY
done
