cd /root/repo
L=deeppreconditioning_b200/lib
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pack or pcg or spmv" 2>&1 | tail -2
run() { echo "== $1 $2"; env $2 DPCG_LIB=$PWD/$L/libdpcg_$1.so timeout 300 python tools/gpu_pack_ab.py --systems 64 --reps 1 $3 2>&1 | grep -v Warning | grep "solves/s\|bitwise\|Error\|error"; }
echo "== default (in-stage windows where they fit)"; timeout 300 python tools/gpu_pack_ab.py --systems 64 --reps 1 2>&1 | grep "solves/s\|bitwise\|rror"
run nowin "" --packed-only
echo "=== trace"; DPCG_LIB=$PWD/$L/libdpcg_tr.so timeout 200 python tools/trace_pipe.py 2>&1 | grep -v Warn | head -20
