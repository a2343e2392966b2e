import sys
sys.path.insert(0, '.')
from ternary_birch_b200 import batch
from ternary_birch_b200.database import EigenvectorDatabase
import tempfile, os
d = tempfile.mkdtemp()
out = {}
with EigenvectorDatabase(os.path.join(d, 'e.sqlite3')) as db:
    for level in range(11, 131):
        if not batch.is_squarefree(level):
            continue
        try:
            g = db.get_genus_with_eigenvectors(level, precise=False)
            out[level] = len(g.rational_eigenvectors())
        except Exception as e:
            out[level] = repr(e)[:60]
print(out)
