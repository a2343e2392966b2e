import sys; sys.path.insert(0,'.')
import torch
from ternary_birch_b200 import Genus, load_genus_table
t = load_genus_table('tests/golden/genus_1062347.npz')
g = Genus(t, precise=False, stream=torch.cuda.current_stream().cuda_stream)
for _ in range(3):
    r = g.hecke_device(101); r.close()
torch.cuda.synchronize()
