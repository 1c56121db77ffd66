#!/bin/bash
python scripts/phase_stamps.py 32768 stoch 2>&1 | tail -3
python scripts/phase_stamps.py 32768 2>&1 | tail -2
BNV_DEBUG_DISABLE=1024 python scripts/phase_stamps.py 32768 stoch 2>&1 | tail -2
