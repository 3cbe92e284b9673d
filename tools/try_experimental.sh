#!/bin/bash
# First GPU call of the next round (ONE gpurun, ~3 minutes): do the two experimental kernel variants work, and what do they buy?
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/try_experimental.sh'
# Both are off by default and have never run (round 1 ended without GPU budget): profiles/README.md, "Backlog".
set -u
mkdir -p gpurun_out
echo "== opt-in comparison tests (variant vs default path, same process layout) =="
WB_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests -m gpu -q -k experimental 2>&1 | tail -15
echo "== bench A/B (ms/step, shade_fwd, decoder_bwd) =="
for cfg in "" "WB_TC_FWD_TMEMA=1" "WB_TC_BWD_GROUPS=3" "WB_TC_FWD_TMEMA=1 WB_TC_BWD_GROUPS=3"; do
  echo "-- ${cfg:-default}"
  env $cfg timeout 250 python bench.py --no-cpu-baseline 2>gpurun_out/try_exp.err | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    print(round(d['ms_per_step'], 2), 'ms/step   fwd', round(d['stage_ms']['shade_fwd'], 3), '  decoder_bwd', round(d['stage_ms']['decoder_bwd'], 3), '  loss', d['e2e']['last_loss'])
except Exception as e:
    print('FAILED:', e); print(open('gpurun_out/try_exp.err').read()[-600:])
"
done
