"""Seeded inputs shared by the golden generator (reference side) and tests/test_goldens.py (oracle / CUDA side)."""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from onnxocr_b200 import synth_pages  # noqa: E402


def det_images():
    rng = np.random.default_rng(100)
    return [rng.integers(0, 256, (97, 131, 3), dtype=np.uint8), rng.integers(0, 256, (20, 30, 3), dtype=np.uint8),
            synth_pages.render_page(300, 1100, 5, seed=101)[0]]


def prob_maps():
    rng = np.random.default_rng(102)
    out = []
    for k in range(3):
        H, W = (160, 240) if k < 2 else (96, 320)
        prob = np.zeros((H, W), np.float32)
        for _ in range(9):
            c = rng.uniform([20, 15], [W - 20, H - 15])
            ang = rng.uniform(-0.4, 0.4)
            bw, bh = rng.uniform(15, 110), rng.uniform(5, 20)
            R = np.array([[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]])
            pts = (np.array([[-bw / 2, -bh / 2], [bw / 2, -bh / 2], [bw / 2, bh / 2], [-bw / 2, bh / 2]]) @ R.T + c).astype(np.int32)
            cv2.fillPoly(prob, [pts], float(rng.uniform(0.5, 1.0)))
        q = np.clip(np.rint(cv2.GaussianBlur(prob, (5, 5), 0) * 255), 0, 255).astype(np.uint8)
        if k == 1:
            for _ in range(5):
                x, y = int(rng.integers(5, W - 8)), int(rng.integers(5, H - 8))
                q[y:y + 3, x:x + 4] = 0
        out.append(((q.astype(np.float32) / np.float32(255)).astype(np.float32), (2 * H + k, 3 * W + k)))
    return out


def crop_case():
    img = synth_pages.render_page(240, 400, 6, seed=103)[0]
    rng = np.random.default_rng(104)
    quads = []
    for t in range(6):
        c = rng.uniform([70, 50], [330, 190])
        ang = rng.uniform(-0.3, 0.3) if t % 2 else 0.0
        bw, bh = rng.uniform(40, 130), rng.uniform(12, 32)
        if t == 5:
            bw, bh = bh, bw
        R = np.array([[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]])
        quads.append((np.array([[-bw / 2, -bh / 2], [bw / 2, -bh / 2], [bw / 2, bh / 2], [-bw / 2, bh / 2]]) @ R.T + c).astype(np.float32))
    return img, quads


def ctc_probs():
    rng = np.random.default_rng(105)
    logits = rng.standard_normal((4, 24, 18385)).astype(np.float32) * 3
    logits[:, ::3, 0] += 30
    logits[3, :, 0] += 100
    p = np.exp(logits - logits.max(axis=2, keepdims=True))
    return (p / p.sum(axis=2, keepdims=True)).astype(np.float32)


def e2e_pages():
    return [synth_pages.render_page(400, 600, 8, seed=106, rot180_frac=0.25)[0],
            synth_pages.render_page(64, 64, 0, seed=107)[0]]
