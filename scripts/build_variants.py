"""Builds tuning variants of the library next to the product (gpboost_b200/lib_var_*.so, git-ignored, shipped to the GPU box by
gpurun) for scripts/time_variants.py:  python scripts/build_variants.py && gpurun -- 'python scripts/time_variants.py
lib_gpboost_b200.so lib_var_nll4.so lib_var_nll6.so lib_var_grad2.so lib_var_grad4.so'
The factor kernel's resident CTAs per SM (register cap = 65536 / (128 threads x CTAs)): the shipped NLL kernel runs 5 CTAs/SM at
96 registers with an 88-byte spill frame (profiles/r01_res_usage.txt); 4 CTAs/SM removes the spills, 6 raises occupancy."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpboost_b200 import build as b

for name, flags in (("lib_var_nll4.so", ["-DGPB_NLL_BLOCKS=4"]), ("lib_var_nll6.so", ["-DGPB_NLL_BLOCKS=6"]),
                    ("lib_var_grad2.so", ["-DGPB_GRAD_BLOCKS=2"]), ("lib_var_grad4.so", ["-DGPB_GRAD_BLOCKS=4"])):
    print(b.build(extra_flags=flags, out_name=name))
