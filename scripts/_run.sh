TAG=x5 scripts/gpu.sh ab 256 tree promo256 promo0
for z in 16 24 32; do MXADV_ZCHUNK=$z TAG=x5z$z scripts/gpu.sh ab 256 tree; done
for z in 32 43; do MXADV_ZCHUNK=$z TAG=x5y$z scripts/gpu.sh ab 512 tree; done
