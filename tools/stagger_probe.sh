for dbg in 0 -1 -2; do BT_TILE_STAGGER_NS=$dbg PROBE_CFGS="11,1,10" python tools/lib_probe.py | sed "s/^/dbg=$dbg /"; done
for tpc in 2 32; do BT_TILE_PER_CTA=$tpc PROBE_CFGS="11,1,10" python tools/lib_probe.py | sed "s/^/tpc=$tpc /"; done
