"""Print a compact table of bench JSON lines (development helper): python tools/bench_table.py files..."""
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    if d.get("impl") == "reference":
        print(f"{f}: reference arm {d['value']:.1f} {d['unit']} ({d['cpu_baseline']['cores']} cores, {d['steps']} steps of {d.get('step_clips')} clips)")
        continue
    r = d["roofline"]; e = d["e2e"]
    print(f"{f}: N={d['n_gpus']} cfg{d['config']['config_id']} value {d['value']:.0f} ({d['ms_per_step']:.3f} ms)  e2e {e['value']:.0f} ({e['ms_per_step']:.3f} ms, "
          f"{e.get('frac_of_copy_ceiling', 0):.2f} of copy ceiling {r.get('copy_ceiling_value', 0):.0f}, bound {e.get('bound')})")
    print(f"     k1 {r['k1_ms']:.3f} k0 {r['k0_ms']:.4f} k2 {r['k2_ms']:.4f} ms  {r['achieved']:.0f} TF/s frac {r['frac']:.3f} (burst {r['frac_of_burst_peak']:.3f}, sustained {r['frac_of_sustained_peak']:.3f}) "
          f"executed {r['executed_frac']:.3f}  traffic {r['traffic']}  k2 hbm {r['k2']['frac']:.3f}  launches {d['gpu_launches']}  clocks {d['clocks']['sm_mhz']} {d['clocks']['reasons']}")
    if "cpu_baseline" in d:
        print(f"     cpu {d['cpu_baseline']['value']:.1f} ({d['cpu_baseline']['cores']} cores)  torch-on-gpu tf32 {d['gpu_torch_baseline'].get('allow_tf32_true',{}).get('value',0):.0f} / fp32 {d['gpu_torch_baseline'].get('allow_tf32_false',{}).get('value',0):.0f}  {d['gpu_torch_baseline'].get('error','')}")
    if "e2e_pcm16" in d:
        print(f"     e2e_pcm16 {d['e2e_pcm16']['value']:.0f}")
