timeout 600 python tools/probe_instep.py 2>&1 | tail -40
