import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cic_b200 as cic
import train_autoencoder as tr
for (B, H, W) in [(3, 64, 64), (2, 64, 64), (3, 32, 48), (4, 128, 128), (3, 64, 128)]:
    model = tr.build_autoencoder((H, W, 3))
    model.set_weights_dict(cic.weights.synthetic_autoencoder(seed=42))
    x = cic.synth.to_unit_range(cic.synth.synth_images_u8(B, H, W, seed=43))
    y = model.predict(x)
    y2 = model.predict(x)
    print((B, H, W), "repeat diff", np.abs(y - y2).max())
    for i in range(B):
        yi = model.predict(x[i:i + 1])[0]
        d = np.abs(yi - y[i])
        bad = np.argwhere(d.max(axis=2) > 1e-6)
        print("  image", i, "max diff", d.max(), "n bad px", len(bad), "rows", sorted(set(bad[:, 0]))[:12], "cols", sorted(set(bad[:, 1]))[:12])
