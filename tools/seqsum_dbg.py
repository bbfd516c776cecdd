import sys, numpy as np
sys.path.insert(0, ".")
import spfresh_b200 as s
ctx = s.Context(0)
v = (np.random.default_rng(1).random(1_000_000, dtype=np.float32) * 200 + 150).astype(np.float32)
ctx.seq_sum_f32(v, 3)
ctx.seq_sum_f32(v, 3)
