"""Generates tests/golden/param_means.json: mean linear RGB of the golden PNGs the reference ships for its 22 parameter-sweep scenes
(doc/2022_q1/img/param/<name>.png <-> scene/2022_q1/parameters/<name>.xml) and of doc/2022_q1/img/WarmStudio.png, box-filtered to
192x108.  Runs only where /root/reference is mounted; the JSON is committed."""
import json, os
import numpy as np
from PIL import Image

REF = "/root/reference"
out = {}
names = sorted(f[:-4] for f in os.listdir(f"{REF}/scene/2022_q1/parameters") if f.endswith(".xml"))
for name in names + ["WarmStudio"]:
    path = f"{REF}/doc/2022_q1/img/param/{name}.png" if name != "WarmStudio" else f"{REF}/doc/2022_q1/img/WarmStudio.png"
    g = np.asarray(Image.open(path).convert("RGB").resize((192, 108), Image.BOX)).astype(np.float64) / 255
    lin = np.where(g <= 0.04045, g / 12.92, ((g + 0.055) / 1.055) ** 2.4)
    out[name] = [round(float(v), 5) for v in lin.mean(axis=(0, 1))]
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "param_means.json"), "w"), indent=1)
print(len(out), "images")
