"""Convert the reference's only map fixture (data/corridor.yaml, 799x199 @0.05 m) to a compact
.npz so the tests can run where /root/reference does not exist (the GPU box).

Run in the build container:  python tests/golden/make_corridor.py
The raw cell values are kept; the reference reader's `data>0 -> 100 else 0` mapping
(test/read_map.hpp:42-45) is applied by tests/fixtures.py at load time.
"""
from pathlib import Path

import numpy as np
import yaml

SRC = Path("/root/reference/data/corridor.yaml")
DST = Path(__file__).resolve().parent / "corridor.npz"

cfg = yaml.safe_load(SRC.read_text())
d = {}
for item in cfg:
    d.update(item)
data = np.asarray(d["data"], dtype=np.int16)
assert data.min() >= -128 and data.max() <= 127
np.savez_compressed(
    DST,
    start=np.asarray(d["start"], dtype=np.float64),
    goal=np.asarray(d["goal"], dtype=np.float64),
    origin=np.asarray(d["origin"], dtype=np.float64),
    dim=np.asarray(d["dim"], dtype=np.int32),
    resolution=np.float64(d["resolution"]),
    data=data.astype(np.int8),
)
print(DST, DST.stat().st_size, "bytes;", (data > 0).sum(), "occupied,", (data < 0).sum(), "unknown")
