for ts in 1664 1792 1920; do echo "== TS=$ts"; PMM_VERBOSE=1 PMM_TILE_SLOTS=$ts bash scripts/quick.sh config4_100k_x_2k 2>&1 | grep -E "pmm\]|estep" | sort -u | head -12; done
