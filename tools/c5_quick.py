"""1-GPU C5 numbers of the large-n section of bench.py (projcg / pcg iterations per second, Gram)."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench, lfpsqp.jl_b200 as L
ctx = L.Context(0); dev = torch.device("cuda", 0); torch.cuda.set_device(0)
o = bench.large_n_section(L, ctx, torch, dev, False)
print("1gpu projcg it/s %.1f frac %.4f | pcg %.1f frac %.4f | gram %.2f ms %.2f TF/s" % (o["value"], o["roofline"]["frac"], o["retraction_pcg"]["value"],
      o["retraction_pcg"]["roofline"]["frac"], o["gram"]["ms"], o["gram"]["achieved"]))
