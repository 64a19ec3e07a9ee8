for mc in 64 1024 4096; do echo "merge_cells=$mc"; CBN_COUNT_MERGE_CELLS=$mc CBN_COUNT_DEBUG=1 python tools/bench_kernels.py count 2>&1 | grep -v "  group"; done
