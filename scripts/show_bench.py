"""prints the interesting parts of a bench.py JSON line"""
import json, sys
lines = [l for l in open(sys.argv[1]) if l.startswith("{")]
print("json lines on stdout:", len(lines))
d = json.loads(lines[-1])
print("value %.1f  e2e %.1f  (S=%d)  single %.1f / e2e %.1f   timed region %.2f s   launches %d" % (d["value"], d["e2e"]["value"], d["config"]["sequences_per_gpu"],
      d["single_sequence"]["value"], d["single_sequence"]["e2e"], d["timed_region_s"], d["gpu_launches"]))
r = d["roofline"]
print("roofline hot %.3f (%.2f us)  cold %.3f (%.2f us)  in tracker %.2f us  traffic %s" % (r["frac"], r["us_per_launch"], r["cold"]["frac"], r["cold"]["us_per_launch"], r["in_tracker"]["us_per_iteration"], r["traffic"]))
print("at 1280x960:", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in (r["at_1280x960"] or {}).items()})
print("cpu_baseline:", d["cpu_baseline"])
print("reference tracker on this GPU:", json.dumps(d.get("reference_tracker_on_this_gpu"))[:900])
print("clocks:", d["clocks"], " ATE:", d["trajectory_ate_rmse_m"])
print("offline io:", {k: v for k, v in d["offline_batch_io"].items() if k != "what"})
for e in (d.get("extra_configs") or []) if isinstance(d.get("extra_configs"), list) else [d.get("extra_configs")]:
    if e and "config" in e:
        print("extra:", e["config"][:34], "fps %.1f" % e["frames_per_s"], "surfels", e["surfels_at_end"], [(c["frame"], c["surfels"]) for c in e["surfels"]], "stages", {k[:14]: round(v, 3) for k, v in e["stage_ms_at_final_map"].items()}, "ATE", e["trajectory_ate_rmse_m"])
    else:
        print("extra:", e)
