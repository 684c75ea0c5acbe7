"""Cosine beta schedule of the UNMODIFIED reference (diffusion/frameworks/utils.py:7-60), float64, T = 1000 and T = 50.

    python tests/golden/make_schedule_golden.py          # needs /root/reference; writes tests/golden/schedule_golden.npz
"""
import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("ref_fw_utils", "/root/reference/diffusion/frameworks/utils.py")
mod = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mod)
np.savez(os.path.join(HERE, "schedule_golden.npz"),
         cosine_1000=mod.get_betas_by_name("cosine", 1000), cosine_50=mod.get_betas_by_name("cosine", 50),
         linear_50=mod.get_betas_by_name("linear", 50))
print("written")
