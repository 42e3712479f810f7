timeout 300 python tools/probe_timeline.py 2>&1 | grep -A6 "^qkv"
MA3_QKV_EW=12 timeout 300 python tools/probe_timeline.py 2>&1 | grep -A6 "^qkv"
timeout 300 python tools/probe_power.py 2>&1 | grep "qkv"
MA3_QKV_EW=12 timeout 300 python tools/probe_power.py 2>&1 | grep "qkv"
