for lp in 0 1; do BT_TILE_LOOP=$lp PROBE_CFGS="12,0,10" python tools/lib_probe.py | sed "s/^/loop=$lp /"; done
