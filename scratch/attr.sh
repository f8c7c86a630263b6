for d in 0 1 2 4 6 7; do echo "debug=$d"; FVFI_CONV_DEBUG=$d python tools/bench_conv.py 2 2>&1 | grep -E "32->32|25->25|64->64 half" | cut -c1-72; done
