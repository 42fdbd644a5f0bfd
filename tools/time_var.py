"""Time the cfg3 forward for each SKB_FWD5_VARIANT in argv (development aid; one process per variant)."""
import os, subprocess, sys
for v in sys.argv[1:]:
    env = dict(os.environ, SKB_FWD5_VARIANT=v, SKB_WPSM="0")
    out = subprocess.run([sys.executable, "tools/time_fwd.py", "cfg3"], env=env, capture_output=True, text=True).stdout
    for ln in out.splitlines():
        print("variant", v, ln, flush=True)
