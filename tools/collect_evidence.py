"""Copy the outputs of tools/evidence_r2.sh from gpurun_out/ into profiles/ (read here, no GPU): ncu summary, traffic
file keyed by the hash of the kernel sources, bench line, launch list, sweeps."""
import json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"),
                    "Round 2 (final sources): ncu --set full of the cfg-2 training step kernels (two-term forward, tile-major backward)",
                    os.path.join(G, "r2_prof.ncu-rep")], capture_output=True, text=True)
open(os.path.join(P, "r2_ncu_full_summary.md"), "w").write(r.stdout)
t = json.loads(r.stderr.strip().splitlines()[-1])["bytes_per_launch"]
json.dump({"bytes_per_launch": t, "src_sha16": bench._src_sha16(),
           "source": "profiles/r2_ncu_full_summary.md (ncu --set full, tools/evidence_r2.sh, cfg-2 training step, two-term forward weights)"},
          open(os.path.join(P, "ncu_traffic.json"), "w"), indent=1)
for a, b in (("r2_launches.csv", "r2_launches.csv"), ("r2_precision_sweep.txt", "r2_precision_sweep.txt")):
    shutil.copy(os.path.join(G, a), os.path.join(P, b))
open(os.path.join(P, "r2_config_sweep.txt"), "w").writelines(l for l in open(os.path.join(G, "r2_sweep.log")) if l.startswith("[cfg"))
line = [l for l in open(os.path.join(G, "r2_bench.json")) if l.startswith("{")][-1]
d = json.loads(line)
if d["roofline"].get("traffic") is not None:  # (a bench line taken before the traffic file matched is not copied)
    shutil.copy(os.path.join(G, "r2_bench.json"), os.path.join(P, "r2_bench.json"))
print("traffic", t, "| bench traffic", d["roofline"].get("traffic"), "| ms/step", d["ms_per_step"])
