python tools/recon_debug.py 2>&1 | tail -40
