#!/bin/bash
# A/B of the table-level fake-quant (default) against fake-quant per gathered corner (PN_QUANT_IN_GATHER=1).
python -m pytest tests -m gpu -x -q -k "quant or acaq or llff or packed or pnq or export" 2>&1 | tail -3
for v in "" 1; do
  PN_QUANT_IN_GATHER=$v python bench.py --workload llff_acaq --no-baselines --no-extras --steps 10 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('in_gather=%r' % '$v', d['ms_per_step'], d['value'], d['roofline']['all_field_launches_ms'] if d.get('roofline') else None)"
done
for v in "" 1; do
  PN_QUANT_IN_GATHER=$v python scripts/bench_render_t22.py 2>/dev/null | tail -1 | cut -c1-600
done
