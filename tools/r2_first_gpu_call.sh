#!/bin/bash
# First GPU call of a new round: confirm everything that was written without a GPU at the end of
# the previous round, then the usual evidence.  Run as
#   gpurun --timeout 1500 -- 'bash tools/r2_first_gpu_call.sh'
# Everything lands in gpurun_out/r2_first/.
set -u
O=gpurun_out/r2_first
mkdir -p $O
# 1. the reference-numbering kernels (non-strict xfail in the suite): run them for real
timeout 900 python -m pytest tests/test_gpu_zz_refnum.py -q -m gpu --runxfail -x > $O/refnum_pytest.log 2>&1
echo "refnum pytest exit $?" >> $O/refnum_pytest.log
# 1a. lx(1) / rx(1) relinked against the CUDA engine
timeout 600 python -m pytest tests/test_gpu_zy_relinked_clis.py -q -m gpu --runxfail > $O/clis_pytest.log 2>&1
echo "relinked CLIs pytest exit $?" >> $O/clis_pytest.log
# 1b. eager outputs (k1_eager.cu, carry in K2/K3, the reference's tests/eager_output programs)
timeout 900 python -m pytest tests/test_gpu_zzz_eager.py -q -m gpu --runxfail > $O/eager_pytest.log 2>&1
echo "eager pytest exit $?" >> $O/eager_pytest.log
timeout 300 python tools/bench_eager.py > $O/bench_eager.json 2> $O/bench_eager.err
# 2. config 5 with both numberings: time split (ms_numbering) + identity with the compiled reference
NUMBERING=bfs timeout 300 python tools/bench_determinise.py > $O/det_bfs.json 2> $O/det_bfs.err
NUMBERING=reference timeout 300 python tools/bench_determinise.py > $O/det_reference.json 2> $O/det_reference.err
# 3. the whole GPU suite (the shim now pins cached tables while a call runs on them)
timeout 1200 python -m pytest tests -q -m gpu -x > $O/gpu_pytest.log 2>&1
echo "gpu pytest exit $?" >> $O/gpu_pytest.log
# 4. thread stress of the shim against the CUDA engine (more threads than cache slots)
timeout 300 build/shim/shim_threads 24 3 > $O/shim_threads.log 2>&1
echo "shim_threads exit $?" >> $O/shim_threads.log
# 5. the bench line
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err
tail -3 $O/clis_pytest.log $O/refnum_pytest.log $O/eager_pytest.log $O/gpu_pytest.log $O/shim_threads.log
cat $O/det_reference.json $O/bench.json
