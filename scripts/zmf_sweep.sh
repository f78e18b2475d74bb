# zernmodfit kernel: frame tiles per warp (ZMF_MT) x split-K (ZMF_KSPLIT) sweep; prints kernel time and HBM roofline fraction
for mt in 1 2; do for ks in 4 8 16 32; do echo "MT=$mt KSPLIT=$ks"; ZMF_MT=$mt ZMF_KSPLIT=$ks python scripts/bench_extra.py zmf 2>&1 | grep zernmodfit | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print('   ', d['workload'][16:30], round(d['kernel_ms']*1e3,1), 'us', round(d['roofline']['frac'],3))
"; done; done
