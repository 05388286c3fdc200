"""Generates tests/golden/interchange/model_<seed>.modal: `.modal` bytes of seeded models written by the UNMODIFIED reference
archive (oracle/_ref, `make -C oracle ref`). The models themselves are regenerated from the seed (oracle/interchange.py).
Run: python tests/golden/make_interchange_golden.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import interchange as oi  # noqa: E402

SEEDS = {3: dict(), 8: dict(n_modes=1, n_points=1, n_eigen=1), 21: dict(n_modes=30, n_points=12, n_eigen=45)}

if __name__ == "__main__":
    for seed, shape in SEEDS.items():
        data = oi.serialize(oi.random_model(seed, **shape))
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "interchange", f"model_{seed}.modal"), "wb") as f:
            f.write(data)
        print(seed, len(data), "bytes")
