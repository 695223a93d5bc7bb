"""One-line summary of a bench.py JSON line (used by the tools/gpu_*.sh scripts)."""
import json
import sys

try:
    l = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r, c = l["roofline"], l["config"]
    print(" value %.3e ms/dem %.4f k_ms %.4f frac %.3f C %.2f T %.2f rebuilds %d (each %.2f ms, share %.3f) e2e %.3e setup %.0fs" % (
        l["value"], c["ms_per_dem_step"], r["kernel_ms"], r["frac"], r["C_half"], r["T_half"], c["rebuilds_in_timed_region"],
        r["rebuild_ms_each"], r["rebuild_share_of_step"], l["e2e"]["value"], c["setup_s"]))
except Exception as e:  # noqa: BLE001
    print(" failed", e)
