#!/bin/bash
# quick single-GPU check: parity suite + one short bench line + phase stamps
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python bench.py --steps 3000 --warmup 50 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value',round(d['value']), 'us/step',round(d['ms_per_step']*1e3,2),'kernel_us',round(d['roofline']['kernel_us'],2),'b2b',round(d['config']['back_to_back_ms_per_step']*1e3,2),'e2e',round(d['e2e']['value']), d['config']['launch'])"
python scripts/phase_stamps.py | tail -2
