cd /root/repo
L=deeppreconditioning_b200/lib
echo "== default"; python tools/time_spmv.py 128 2>&1 | grep "\^3"
for v in s4 rr2s4 rr4s4 rr8s4 rr4; do echo "== $v"; DPCG_LIB=$PWD/$L/libdpcg_$v.so python tools/time_spmv.py 128 2>&1 | grep "\^3"; done
