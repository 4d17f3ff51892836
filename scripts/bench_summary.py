import json, sys
for f in sys.argv[1:]:
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    r = d["roofline"]
    print("%s: %.1f evals/s %.3f ms/step e2e %.3f ms | eval frac %.3f | %s" % (
        f, d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], r.get("eval_frac_of_peak_all_gpus", 0),
        {k: v for k, v in r["families_ms"].items() if v}))
    if "n200k" in d:
        n = d["n200k"]; r = n["roofline"]
        print("   n200k: %.2f evals/s %.2f ms/step | eval frac %.3f | %s" % (
            n["value"], n["ms_per_step"], r.get("eval_frac_of_peak_all_gpus", 0),
            {k: v for k, v in r["families_ms"].items() if v}))
