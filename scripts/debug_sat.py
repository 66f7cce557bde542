import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import regularizepsf_b200 as rp
from oracle import cpu_oracle as oracle
from tests.helpers import load_golden
import warnings
warnings.simplefilter("ignore")
g = load_golden("p32_saturation")
coords, size = g["coords"], g["size"]
kernel = np.ones((len(coords), size, size), dtype=np.complex128)
image = g["image"]
kw = g["apply_kwargs"]
want = oracle.apply_transform(image, coords, kernel, **kw)
t = rp.ArrayPSFTransform(rp.IndexedCube(coords, kernel))
got = t.apply(image, dtype="float64", **kw)
d = np.abs(got - want)
print("max err", d.max(), "n masked", int((image > kw["saturation_threshold"]).sum()))
idx = np.argsort(d.ravel())[::-1][:12]
for k in idx:
    y, x = divmod(int(k), image.shape[1])
    print((y, x), "got", got[y, x], "want", want[y, x], "img", image[y, x])
# direct comparison of fills: emulate reference fill on padded
p = size
padded = np.pad(image.astype(float), 2 * p, mode="symmetric")
mask = oracle.fill_saturated(padded, kw["saturation_threshold"], kw.get("saturation_dilation", 1), kw.get("neighborhood_width", 7))
crop = padded[2 * p:-2 * p, 2 * p:-2 * p]
m = mask[2 * p:-2 * p, 2 * p:-2 * p]
print("identity check (oracle out vs its own filled frame, unmasked):", np.abs(want - crop)[~m].max())
print("ours vs filled frame, unmasked:", np.abs(got - crop)[~m].max(), " masked count in crop", int(m.sum()))
ys, xs = np.where(m)
print("masked rows range", ys.min(), ys.max(), "cols", xs.min(), xs.max())
print("---- real kernel")
from tests.test_gpu_parity import oracle_kernel
kernel = oracle_kernel(g)
print("alpha eps", g["alpha"], g["epsilon"], "|K|max", np.abs(kernel).max())
want = g["out"]
t = rp.ArrayPSFTransform(rp.IndexedCube(coords, kernel))
for dt in ("float32", "float64"):
    got = t.apply(image, dtype=dt, **kw)
    d = np.abs(got - want)
    print(dt, "max err", d.max(), "rel", d.max() / image.max())
    y, x = np.unravel_index(np.argmax(d), d.shape)
    print("  at", (y, x), got[y, x], want[y, x])
    # feed the oracle-filled frame without saturation
    got2 = t.apply(crop.astype(np.float32) if dt == "float32" else crop, dtype=dt)
    # compare away from masked pixels
    want2 = oracle.apply_transform(crop, coords, kernel)
    print("  no-sat path on filled frame: err", np.abs(got2 - want2).max())
    print("  want(sat) vs want2 at unmasked:", np.abs(want - want2)[~m].max())
