PUMAS_B200_LIBRARY=$PWD/variants/c2.so timeout 200 python tests/_gpu_kernels.py -k 0 headline:600000 straight:1000000 2>&1 | tail -2
PUMAS_B200_LIBRARY=$PWD/variants/c2.so timeout 200 python tests/_gpu_kernels.py -k 0,1 headline:30000 straight_mixed:30000 straight:30000 2>&1 | grep bit
