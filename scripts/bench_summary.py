import json, sys
for f in sys.argv[1:]:
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    r = j["roofline"]
    print(f"{f}: value {j['value']:.0f} ({j['ms_per_step']:.2f} ms/step) e2e {j['e2e']['value']:.0f} gemm {r['achieved']:.0f} TF ({r['frac']:.3f}) whole {r['whole_step_frac_of_peak']:.3f} sm_mhz {j['clocks'].get('sm_mhz')} {j['clocks'].get('reasons')}")
    print("   shares", {k: round(v, 3) for k, v in r["share_of_step"].items()}, "mel GB/s", round(r["mel_stage"]["achieved_gbs"]), "launches", j["gpu_launches"])
