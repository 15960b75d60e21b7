#!/bin/bash
# A/B timing of the tile kernel across builds of the same ABI (HGP_LIB selects the library)
for lib in "$@"; do HGP_LIB=$PWD/hdpgpc_b200/lib/$lib python tools/tile_bench.py 100000 5 2>&1 | tail -1; done
