import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(f"value {d['value']:.0f} {d['unit']}  {d['ms_per_step']:.3f} ms/step  e2e {d['e2e']['value']:.0f}  launches/step {d.get('launches_per_step')}  clocks {d.get('clocks')}")
r = d["roofline"]
print(f"roofline: {r['kernel']} bound={r['bound']} achieved={r['achieved']:.1f} {r['unit']} frac={r['frac']:.3f} share={r['share_of_step']:.3f}  step_hbm_frac={r['step_hbm_frac']:.4f} step_tf32_frac={r['step_tf32_frac']:.4f}")
for k in r["kernels"]:
    print(f"  {k['name']:18s} x{k['launches']:3d} {k['ms_per_step']:8.3f} ms  {k['GBps'] or 0:8.1f} GB/s {k['TFLOPs'] or 0:7.2f} TF/s")
if d.get("cpu_baseline"):
    print("cpu_baseline", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
for k, v in (d.get("other_configs") or {}).items():
    if "error" in v:
        print(f"  {k}: FAILED {v['error']}")
        continue
    print(f"  {k}: {v['samples_s']:.0f} samples/s {v['ms_per_step']:.3f} ms/step hbm_frac {v['step_hbm_frac']:.3f} e2e {v['e2e']['value']:.0f} launches {v.get('launches_per_step')} graph {v.get('cuda_graph')} eager {v.get('eager_ms_per_step')}")
for k in ("e2e_raw", "step_api", "eager", "replica_check", "replica", "ddp_leg"):
    if k in d:
        print(k, d[k])
