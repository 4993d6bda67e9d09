import numpy as np, sys
def gen(path, rows, seed):
    rng = np.random.default_rng(seed)
    w = rng.normal(0, 1, 400)
    with open(path, "w") as f:
        for _ in range(rows):
            ids = np.unique(rng.integers(1, 400, rng.integers(5, 25)))
            x = rng.random(len(ids)).astype(np.float32)
            y = 1 if (w[ids] * x).sum() + rng.normal(0, 0.3) > 0 else -1
            f.write(f"{y} " + " ".join(f"{int(i) * 7919}:{float(v):.6g}" for i, v in zip(ids, x)) + "\n")
gen(sys.argv[1], 4000, 1)
