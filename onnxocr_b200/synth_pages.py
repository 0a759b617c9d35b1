"""Synthetic document pages / text-line crops for BASELINE.json configs C2–C5.

No dataset can be fetched (no network) and `onnxocr/fonts/simfang.ttf` is missing
from the reference mount, so glyphs come from PIL's built-in scalable font; the
generators pick up `simfang.ttf` automatically if `font_path` is given.  Layout is
driven by `numpy.random.default_rng(seed)` only, so every box renders the same bytes.
"""
import numpy as np
from PIL import Image, ImageDraw, ImageFont

_WORDS = ("invoice total amount due date number street city order item price tax net gross "
          "B200 tensor memory page line OCR PP-OCRv5 2026 10 18 0.45 1,250.00 ref id "
          "customer payment bank account signature approved shipped received").split()


def _font(size, font_path=None):
    if font_path:
        return ImageFont.truetype(font_path, size)
    return ImageFont.load_default(size=size)


def _sentence(rng, n_chars):
    out = []
    ln = 0
    while ln < n_chars:
        w = _WORDS[int(rng.integers(len(_WORDS)))]
        out.append(w)
        ln += len(w) + 1
    return " ".join(out)[:n_chars].strip() or "text"


def render_page(h, w, n_lines, seed, rot180_frac=0.0, font_path=None, min_px=20, max_px=32, columns=1):
    """White page (BGR u8, [h,w,3]) with up to `n_lines` dark text lines in `columns` columns; returns (img, line_boxes)."""
    rng = np.random.default_rng(seed)
    page = Image.new("RGB", (w, h), (255, 255, 255))
    draw = ImageDraw.Draw(page)
    margin = max(8, w // 40)
    boxes = []
    per_col = (n_lines + columns - 1) // columns
    col_w = (w - margin) // columns
    gap_budget = max(0, h - 2 * margin - per_col * ((max_px + min_px) // 2 + 4))
    for col in range(columns):
        x0 = margin + col * col_w
        x1 = x0 + col_w - margin
        y = margin
        for i in range(per_col):
            if len(boxes) >= n_lines:
                break
            px = int(rng.integers(min_px, max_px + 1))
            if y + px + 4 >= h - margin:
                break
            font = _font(px, font_path)
            n_chars = int(rng.integers(8, max(9, int((x1 - x0) / (0.55 * px)))))
            text = _sentence(rng, n_chars)
            x = x0 + int(rng.integers(0, max(1, col_w // 10)))
            shade = int(rng.integers(0, 80))
            bbox = draw.textbbox((x, y), text, font=font)
            if bbox[2] > x1:
                keep = max(3, int(len(text) * (x1 - x) / max(1, bbox[2] - x)) - 1)
                text = text[:keep].strip() or "ab"
                bbox = draw.textbbox((x, y), text, font=font)
            if rot180_frac > 0 and rng.random() < rot180_frac:
                tw, th = bbox[2] - bbox[0] + 4, bbox[3] - bbox[1] + 4
                tile = Image.new("RGB", (tw, th), (255, 255, 255))
                ImageDraw.Draw(tile).text((2 - (bbox[0] - x), 2 - (bbox[1] - y)), text, font=font, fill=(shade,) * 3)
                page.paste(tile.rotate(180), (bbox[0] - 2, bbox[1] - 2))
            else:
                draw.text((x, y), text, font=font, fill=(shade, shade, shade))
            boxes.append(bbox)
            y = bbox[3] + 4 + int(rng.integers(0, max(1, 2 * gap_budget // max(1, per_col) + 1)))
    img = np.ascontiguousarray(np.array(page)[:, :, ::-1])
    return img, boxes


def det_pages_c2(n=16, side=960, lines=40, seed=0, font_path=None):
    """C2: `n` pages of side×side×3 u8 BGR, `lines` text lines each."""
    return [render_page(side, side, lines, seed * 1000 + i, font_path=font_path, columns=2)[0] for i in range(n)]


def rec_crops_c3(n=256, seed=1, h=48, w=320, font_path=None):
    """C3: `n` crops 48×320×3 u8 BGR, one rendered text line each (10–25 chars)."""
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        img = Image.new("RGB", (w, h), (255, 255, 255))
        text = _sentence(rng, int(rng.integers(10, 26)))
        px = int(rng.integers(24, 36))
        ImageDraw.Draw(img).text((4, (h - px) // 2 - 2), text, font=_font(px, font_path),
                                 fill=(int(rng.integers(0, 80)),) * 3)
        out.append(np.ascontiguousarray(np.array(img)[:, :, ::-1]))
    return out


_C4_SIZES = [(720, 1280), (1080, 1920), (1754, 1240), (2339, 1654), (960, 960)]


def e2e_pages_c4(n=64, seed=2, font_path=None):
    """C4: mixed-size pages, 20–80 lines, 10 % of lines rotated 180°."""
    rng = np.random.default_rng(seed)
    pages = []
    for i in range(n):
        h, w = _C4_SIZES[int(rng.integers(len(_C4_SIZES)))]
        lines = int(rng.integers(20, 81))
        pages.append(render_page(h, w, lines, seed * 1000 + i, rot180_frac=0.1, font_path=font_path,
                                 min_px=max(16, h // 60), max_px=max(24, h // 40))[0])
    return pages


def pdf_pages_c5(n=8, seed=3, font_path=None):
    """C5: A4 @ 300 dpi pages (3508×2480×3 u8), as `pdf_to_images` would yield them."""
    return [render_page(3508, 2480, 60, seed * 1000 + i, font_path=font_path, min_px=40, max_px=64)[0]
            for i in range(n)]
