timeout 900 python profiles/sweep.py 20 duo:160 pipe:160 duo:180 pipe:180 duo:444 pipe:444 wpc:444 duo:592 wpc:592 duo:740 wpc:740 duo:888 wpc:888 duo:1036 wpc:1036 duo:1184 wpc:1184 wpc:2220 2>&1 | tail -18
timeout 300 python profiles/sweep.py 16 wpc:2220 2>&1 | tail -1
