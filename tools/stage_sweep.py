import os, sys, shutil, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "reni_b200/lib/libreni_b200.so")
shutil.copy(lib, lib + ".bak")
for n in ("2", "3", "4"):
    src = lib + ".bak" if n == "4" else os.path.join(ROOT, f"reni_b200/lib/libreni_stages{n}.so")
    shutil.copy(src, lib)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools/gpu_check.py"), "timing"], capture_output=True, text=True).stdout
    print("stages", n); print("\n".join(l for l in out.splitlines() if "timing" in l))
shutil.copy(lib + ".bak", lib)
