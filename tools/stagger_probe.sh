for ns in 0 -1 -2; do BT_TILE_STAGGER_NS=$ns PROBE_CFGS="12,0,8" python tools/lib_probe.py | sed "s/^/dbg=$ns /"; done
