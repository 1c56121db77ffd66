#!/usr/bin/env python
"""One-line digest of a bench.py JSON line read from stdin (session scripts)."""
import json
import sys

for raw in sys.stdin:
    raw = raw.strip()
    if not raw.startswith("{"):
        continue
    d = json.loads(raw)
    if "unavailable" in d:
        print(d)
        continue
    det, roof, e2e = d.get("detail", {}), d.get("roofline", {}), d.get("e2e", {})
    parts = [d["config"]["name"], f"N={d['n_gpus']}", f"value={d['value']:.1f}", f"us/step={d['ms_per_step'] * 1e3:.2f}"]
    if roof:
        parts += [f"kernel_us={roof['kernel_us']:.2f}", f"GB/s={roof['achieved']:.0f}", f"frac={roof['frac']:.3f}"]
    if "back_to_back_ms_per_step" in det:
        parts.append(f"b2b_us={det['back_to_back_ms_per_step'] * 1e3:.2f}")
    if e2e:
        parts.append(f"e2e={e2e['value']:.1f} ({e2e.get('mode', e2e.get('note', ''))[:24]})")
        for k in ("value_plain_launch", "value_prelaunched", "value_forward_then_cpu"):
            if k in e2e:
                parts.append(f"{k[6:]}={e2e[k]:.0f}")
    if "launch" in det:
        parts.append(str(det["launch"]))
    if "parity_check" in d:
        parts.append(f"parity={d['parity_check']} du={det['parity']['max_abs_du_vs_unsharded']:.1e}")
    if "speedup_vs_1gpu" in det:
        parts.append(f"1gpu_same_total_us={det['single_gpu_same_total_ms'] * 1e3:.1f} speedup={det['speedup_vs_1gpu']:.2f}")
    if "cpu_baseline" in d:
        parts.append(f"cpu={d['cpu_baseline']['value']:.2f}/{d['cpu_baseline']['cores']}c")
    print(" ".join(parts))
