#!/bin/bash
# One GPU-box session: parity tests, then the bench (+ optional A/B variants given as "NAME:ENV=VAL,ENV=VAL" args).
# Everything is logged under gpurun_out/ (merged back into the repo by gpurun).
mkdir -p gpurun_out
TAG=${TAG:-run}
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q ${PYTEST_ARGS:-} > gpurun_out/${TAG}_pytest.log 2>&1
  echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
  tail -5 gpurun_out/${TAG}_pytest.log
fi
timeout 600 python bench.py --steps ${STEPS:-10} --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"; cat gpurun_out/${TAG}_bench.json | python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'hostrays', round(d['e2e']['host_rays_variant']['ms_per_step'],3), {k: round(v,3) for k,v in d['roofline']['stage_ms'].items()}, 'frac', round(d['roofline']['frac'],3))
"
for v in "$@"; do
  name=${v%%:*}; envs=${v#*:}
  ( IFS=,; for e in $envs; do export "$e"; done
    timeout 600 python bench.py --steps ${STEPS:-10} --warmup 3 > gpurun_out/${TAG}_bench_${name}.json 2> gpurun_out/${TAG}_bench_${name}.err
    echo "variant $name exit $?"; python -c "
import json,sys
for l in open('gpurun_out/${TAG}_bench_${name}.json'):
    try: d=json.loads(l)
    except Exception: continue
    print('$name ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), {k: round(v,3) for k,v in d['roofline']['stage_ms'].items()})
" )
done
