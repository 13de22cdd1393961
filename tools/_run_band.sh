python -m pytest tests/test_gpu_kernels.py -q -x -k "band" 2>&1 | tail -2
for a in -1 0 2 4 8 16; do echo "ahead $a"; EGP_BAND_RUN_AHEAD=$a python tools/probe_band_wide.py 16 | cut -c1-90; done
EGP_BAND_RUN_AHEAD=4 python tools/probe_band_wide.py 8 | cut -c1-90
EGP_BAND_RUN_AHEAD=4 EGP_BAND_RUN_CTAS=2 python tools/probe_band_wide.py 16 | cut -c1-90
