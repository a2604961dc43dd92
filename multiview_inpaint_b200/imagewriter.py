"""Asynchronous PNG / NPY writer for the inference loops.

The reference's render scripts call render() and then block on `torchvision.utils.save_image` / `np.save` for every
view (gs-simp/render.py:32-39, render_depth.py:31-39, gen_seq.py:36-58, vis_render.py:37-51): a synchronous
device-to-host copy of a float image, an 8-bit conversion on the CPU and a PNG encode, all on the thread that should
be queueing the next view.  Here

  * the float -> 8-bit conversion runs on the GPU (gsr_quantize_rgb8, bit-identical to save_image's
    `mul(255).add_(0.5).clamp_(0, 255).to(uint8)`), on the stream that rendered the view;
  * the 3-bytes-per-pixel result is copied into a ring of PINNED host buffers with a non-blocking copy and an event;
  * worker threads wait for the event, encode (the zlib encoder below, or PIL -- what torchvision uses -- with
    use_pil=True; the decoded pixels are identical either way) and write the file.  zlib and file I/O release the GIL.

`submit_png` only blocks when every ring slot is still being drained (back-pressure instead of unbounded memory).
Use as the `sink` of multiview.cuda_views_render:

    with AsyncImageWriter(device, slots=8, workers=4) as w:
        cuda_views_render(g, settings, pipeline=pipe, workspaces=ws,
                          sink=lambda k, color, depth, radii: w.submit_png(f"{out}/{k:05d}.png", color))
"""
from __future__ import annotations

import os
import queue
import struct
import threading
import zlib

import numpy as np
import torch

from . import _C

try:  # torchvision.utils.save_image ends in PIL.Image.fromarray(ndarr).save(fp)
    from PIL import Image as _PILImage
except Exception:  # pragma: no cover
    _PILImage = None


def encode_png_rgb8(rgb: np.ndarray, level: int = 3) -> bytes:
    """Minimal PNG encoder (8-bit RGB, filter type 0 on every row, one IDAT): (H,W,3) uint8 -> file bytes."""
    assert rgb.dtype == np.uint8 and rgb.ndim == 3 and rgb.shape[2] == 3
    H, W, _ = rgb.shape
    raw = np.empty((H, 1 + 3 * W), dtype=np.uint8)
    raw[:, 0] = 0
    raw[:, 1:] = rgb.reshape(H, 3 * W)

    def chunk(tag: bytes, data: bytes) -> bytes:
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)
    ihdr = struct.pack(">IIBBBBB", W, H, 8, 2, 0, 0, 0)
    return b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", ihdr) + chunk(b"IDAT", zlib.compress(raw.tobytes(), level)) + chunk(b"IEND", b"")


class _Slot:
    def __init__(self):
        self.dev = None      # device staging (uint8 / raw bytes)
        self.host = None     # pinned
        self.event = torch.cuda.Event()


class AsyncImageWriter:
    def __init__(self, device, slots: int = 8, workers: int = 4, use_pil: bool = False, png_level: int = 1):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("AsyncImageWriter needs a CUDA device (the 8-bit conversion runs in gsr_quantize_rgb8)")
        # default: the zlib encoder below at level 1 (about 2x faster than PIL's adaptive-filter encoder on a 1080p
        # render, same decoded pixels); use_pil=True writes through PIL like torchvision does
        self.use_pil = bool(use_pil) and _PILImage is not None
        self.png_level = png_level
        self._free: queue.Queue = queue.Queue()
        for _ in range(max(slots, 1)):
            self._free.put(_Slot())
        self._jobs: queue.Queue = queue.Queue()
        self._errors: list = []
        self._pending = 0
        self._cv = threading.Condition()
        self._threads = [threading.Thread(target=self._run, daemon=True) for _ in range(max(workers, 1))]
        for t in self._threads:
            t.start()
        self.bytes_d2h = 0

    # ---- producer side (the thread that queues GPU work) ----
    def _stage(self, nbytes: int) -> _Slot:
        slot = self._free.get()                      # blocks only when every slot is still being drained
        if slot.host is None or slot.host.numel() < nbytes:
            slot.host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
            slot.dev = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        return slot

    def _enqueue(self, slot, job):
        slot.event.record(torch.cuda.current_stream(self.device))
        with self._cv:
            self._pending += 1
        self._jobs.put((slot, job))

    def submit_png(self, path: str, image: torch.Tensor, affine: torch.Tensor | None = None):
        """image (3,H,W) or (1,H,W) float32 on the device, values as save_image expects them (0..1).  Runs on the
        CURRENT stream; returns as soon as the conversion and the copy are queued."""
        C, H, W = image.shape
        n = H * W * 3
        slot = self._stage(n)
        with torch.cuda.device(self.device):
            _C.quantize_rgb8(image if image.is_contiguous() else image.contiguous(), out=slot.dev[:n].view(H, W, 3), affine=affine)
            slot.host[:n].copy_(slot.dev[:n], non_blocking=True)
        self.bytes_d2h += n
        self._enqueue(slot, ("png", path, (H, W, 3), n))

    def submit_npy(self, path: str, tensor: torch.Tensor):
        """np.save(path, tensor.cpu().numpy()) without blocking (depth maps, poses: gen_seq.py:60-61)."""
        t = tensor.contiguous()
        n = t.numel() * t.element_size()
        slot = self._stage(n)
        with torch.cuda.device(self.device):
            slot.dev[:n].copy_(t.view(-1).view(torch.uint8))
            slot.host[:n].copy_(slot.dev[:n], non_blocking=True)
        self.bytes_d2h += n
        self._enqueue(slot, ("npy", path, tuple(t.shape), n, str(t.dtype).replace("torch.", "")))

    # ---- consumer side ----
    def _run(self):
        while True:
            item = self._jobs.get()
            if item is None:
                return
            slot, job = item
            try:
                slot.event.synchronize()
                kind, path = job[0], job[1]
                if kind == "png":
                    _, _, shape, n = job
                    arr = slot.host[:n].numpy().reshape(shape)
                    if self.use_pil:
                        _PILImage.fromarray(arr).save(path, format="PNG")
                    else:
                        with open(path, "wb") as f:
                            f.write(encode_png_rgb8(arr, self.png_level))
                else:
                    _, _, shape, n, dtype = job
                    arr = slot.host[:n].numpy().view(np.dtype(dtype)).reshape(shape)
                    with open(path if path.endswith(".npy") else path + ".npy", "wb") as f:
                        np.save(f, arr)
            except Exception as ex:  # surfaced by flush()
                self._errors.append((job[1], repr(ex)))
            finally:
                self._free.put(slot)
                with self._cv:
                    self._pending -= 1
                    self._cv.notify_all()

    def flush(self):
        """Waits until every submitted file is on disk; raises if any write failed."""
        with self._cv:
            while self._pending:
                self._cv.wait()
        if self._errors:
            errs, self._errors = self._errors, []
            raise RuntimeError(f"AsyncImageWriter: {len(errs)} write(s) failed, first: {errs[0]}")

    def close(self):
        self.flush()
        for _ in self._threads:
            self._jobs.put(None)
        for t in self._threads:
            t.join()
        self._threads = []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

