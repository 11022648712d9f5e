S="qkv m8192,geglu m8192,qkv m131072,proj_in m8192"
run() { name=$1; shift; env "$@" timeout 120 python scripts/gemm_lab.py $name "$S" 2>&1 | grep -v "^$"; env "$@" CA_GEMM_TIMING=1 timeout 120 python scripts/gemm_lab.py $name "$S" 2>&1 | grep "timing\]" | awk 'NR%2==0' | sed 's/\[ca_linear timing\] //'; }
run A CA_GEMM_CFG=256,1,0
run B CA_GEMM_CFG=256,1,0 CA_GEMM_ROLES_LOW=1
run C CA_GEMM_CFG=256,1,0 CA_GEMM_STAGES=6
run D CA_GEMM_CFG=256,1,0 CA_GEMM_VEC32=0
run E CA_GEMM_CFG=240,1,0
run F CA_GEMM_CFG=240,1,0 CA_GEMM_ROLES_LOW=1
