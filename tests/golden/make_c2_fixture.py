"""Writes tests/golden/c2_soa.npz: the SoA export (layout of lite_tess_soa) of the default
bench input, config C2 — lite_b200.generate_crtt(10000, seed=5), 10 810 cells.  bench.py's
`--impl reference` arm loads it with numpy so that the CPU arm never touches the product
library; tests/test_host_model_cpu.py checks the file is what the generator produces."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import lite_b200  # noqa: E402

t, _tau = lite_b200.generate_crtt(10000, seed=5)
soa = t.export()
np.savez_compressed(os.path.join(HERE, "c2_soa.npz"), int_length=np.float64(soa.int_length), **soa.a)
print("cells", soa.n_cells(), "bytes", os.path.getsize(os.path.join(HERE, "c2_soa.npz")))
