"""Print the interesting fields of bench.py JSON lines: python show_bench.py file..."""
import json
import sys

for f in sys.argv[1:]:
    l = [x for x in open(f) if x.startswith("{")]
    if not l:
        print(f, "NO JSON:", open(f).read()[-600:])
        continue
    d = json.loads(l[-1])
    print(f, "gpus", d["n_gpus"], "value %.4g" % d["value"], "ms/step %.1f" % d["ms_per_step"],
          "it/s %.0f" % d.get("pcg_iters_per_s", 0),
          {k: round(v, 2) for k, v in d.get("stage_ms", {}).items()},
          "e2e %.4g" % d["e2e"]["value"])
    k = d["roofline"].get("kernels", {})
    print("   ", {n: (round(v["ms"], 3), round(v["GBps"]), round(v["frac"], 3)) for n, v in k.items()},
          "tol:", d.get("to_tolerance") and (d["to_tolerance"]["iters"], d["to_tolerance"]["ms_step"],
                                              d["to_tolerance"]["max_abs_div"]))
