"""File formats of the reference (SURVEY.md Annex C.3), host side only.

  raw input      header-less W*H bytes per frame at offset frame*W*H   (IO/IOManager.ipp:72-112)
  codestream     frames of uint16 little-endian words, appended        (IO/IOManager.ipp:176-190,615-620)
  <out>_SIZE     ASCII decimal frame lengths (uint16 units) joined by ',' without newline (:184-189)
  decoded still  PGM "P5\\n<W> <H>\\n<2^bd-1>\\n" + W*H bytes             (:267-344)
  decoded video  raw W*H bytes per frame, appended                     (:214-231)
"""
from __future__ import annotations

import os

import numpy as np


def read_raw_frames(path: str, width: int, height: int, first: int, count: int) -> np.ndarray:
    n = width * height
    with open(path, "rb") as f:
        f.seek(first * n)
        buf = f.read(count * n)
    if len(buf) != count * n:
        raise IOError(f"{path}: wanted {count} frames of {width}x{height} from frame {first}, file too short")
    return np.frombuffer(buf, dtype=np.uint8).reshape(count, height, width)


def read_pgm_payload(path: str, width: int, height: int) -> np.ndarray:
    """P5 reader following the reference's loadImageChar rule: header = filesize - W*H bytes
    (IO/IOManager.ipp:36-65)."""
    data = np.fromfile(path, dtype=np.uint8)
    n = width * height
    if data.size < n:
        raise IOError(f"{path}: smaller than {width}x{height}")
    return data[data.size - n:].reshape(height, width)


def write_pgm(path: str, img: np.ndarray, bit_depth: int = 8) -> None:
    h, w = img.shape
    with open(path, "wb") as f:
        f.write(f"P5\n{w} {h}\n{(2 << (bit_depth - 1)) - 1}\n".encode("ascii"))
        f.write(np.ascontiguousarray(img, dtype=np.uint8).tobytes())


def write_size_sidecar(path: str, lengths) -> None:
    with open(path + "_SIZE", "w") as f:
        f.write(",".join(str(int(v)) for v in lengths))


def read_size_sidecar(path: str, frames: int | None = None) -> np.ndarray:
    txt = open(path + "_SIZE").read().strip()
    vals = [int(t) for t in txt.split(",") if t != ""]
    if frames is not None:
        vals = vals[:frames]
    return np.asarray(vals, dtype=np.int32)


def remove_existing(path: str) -> None:
    """replaceExistingFile (IO/IOManager.ipp:158-170): output and sidecar are deleted first."""
    for p in (path, path + "_SIZE"):
        if os.path.exists(p):
            os.remove(p)
