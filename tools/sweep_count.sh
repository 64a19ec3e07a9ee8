echo auto; CBN_COUNT_DEBUG=1 python tools/bench_kernels.py count 2>&1 | grep -v "  group"
echo "tpb 256"; CBN_COUNT_TPB=256 CBN_COUNT_DEBUG=1 python tools/bench_kernels.py count 2>&1 | grep -v "  group"
echo "tpb 512"; CBN_COUNT_TPB=512 CBN_COUNT_DEBUG=1 python tools/bench_kernels.py count 2>&1 | grep -v "  group"
