"""Frame producer / consumer either side of the render path.

The reference builds every test frame on the host -- ``AnimationDataset.__getitem__`` (datasets/animation.py:163-206)
makes the ray tensors with numpy, reloads and resizes the HDRI with cv2 for *every* frame, a DataLoader worker
pickles ~35 MB per frame to the main process, Lightning copies it to the GPU -- and converts every output image to
uint8 on the host (utils/mixins.py:43-58).  Once a frame renders in under a second that host work dominates.
Here the per-frame inputs are produced on the device (``ia_make_rays``), the envmap is uploaded once, and the
images are clipped / scaled to uint8 on the device (``ia_pack_rgb8``) so a frame leaves as 1 byte per channel.

Same batch contract as the reference's ``preprocess_data`` (systems/intrinsic_avatar.py:84-116): ``rays [H*W,8]``,
``body_pose [1,69]``, ``global_orient [1,3]``, ``transl [1,3]``, ``betas``, ``hdri [1024,2048,3]``, ``index``.
"""
from __future__ import annotations

import base64
import os
import queue
import struct
import threading
import zlib

import numpy as np
import torch


class AnimationFrames:
    """Device-side stand-in for AnimationDataset(split="test") + preprocess_data.

    poses [F,72] (global_orient + body_pose), trans [F,3]: ``load/animation/<seq>/poses.npz``;
    K [3,3]: cameras.npz intrinsic already divided by ``downscale``; w2c [4,4] or [F,4,4]: cameras.npz extrinsic.
    """

    def __init__(self, engine, poses, trans, K, H, W, w2c=None, hdri=None, betas=None, near=None, far=None,
                 start=0, end=None, skip=1):
        self.engine = engine
        poses = np.asarray(poses, np.float32)
        trans = np.asarray(trans, np.float32)
        # datasets/animation.py:127-139: the FULL sequence is re-based so that its frame 0 stands at (0, 0.15, 5), and
        # only then is start:end:skip applied -- pass the whole poses.npz and the config's start / end / skip here
        trans = trans - trans[0] + np.array([0, 0.15, 5], np.float32)
        sl = slice(int(start), None if end is None else int(end), int(skip))
        self.poses, self.trans = poses[sl], trans[sl]
        self.K, self.H, self.W = np.asarray(K, np.float64), int(H), int(W)
        self.w2c = None if w2c is None else np.asarray(w2c, np.float32)
        if self.w2c is not None and self.w2c.ndim == 3:
            self.w2c = self.w2c[sl]
        self.near, self.far = near, far
        self.betas = np.zeros((1, 10), np.float32) if betas is None else np.asarray(betas, np.float32).reshape(1, 10)
        # uploaded once; the reference re-reads and re-sizes the .hdr file for every frame (:191-201)
        self.hdri = None if hdri is None else torch.as_tensor(hdri, dtype=torch.float32).to(engine.dev).contiguous()
        self._rays = torch.empty(self.H * self.W, 8, device=engine.dev)

    def __len__(self):
        return len(self.poses)

    def __getitem__(self, idx):
        transl = self.trans[idx]
        if self.near is not None and self.far is not None:
            near, far = self.near, self.far
        else:
            dist = float(np.sqrt(np.square(transl).sum(-1)))      # distance from the camera to the mid-hip (:185-189)
            near, far = dist - 1, dist + 1
        w2c = None
        if self.w2c is not None:
            w2c = self.w2c[idx] if self.w2c.ndim == 3 else self.w2c
        rays = self.engine.make_rays(self.K, self.H, self.W, near, far, w2c=w2c, out=self._rays)
        batch = {
            "rays": rays,
            "betas": torch.from_numpy(self.betas[0]),
            "global_orient": torch.from_numpy(self.poses[idx, :3][None]),
            "body_pose": torch.from_numpy(self.poses[idx, 3:][None]),
            "transl": torch.from_numpy(transl[None]),
            "index": idx,
        }
        if w2c is not None:
            batch["w2c"] = torch.from_numpy(np.asarray(w2c))
        if self.hdri is not None:
            batch["hdri"] = self.hdri
        return batch

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]


def images_to_uint8(engine, out: dict, H: int, W: int, keys=("comp_rgb_full", "comp_rgb_phys_full", "comp_albedo_full"),
                    bgr=True) -> dict:
    """The validation/test image columns of the reference (systems/intrinsic_avatar.py:436-530 through
    SaverMixin.get_rgb_image_) as uint8 [H,W,3] host arrays ready for cv2.imwrite: conversion on the device, one
    3-byte-per-pixel copy per image."""
    res = {}
    for k in keys:
        img = out[k]
        if not img.is_cuda:
            img = img.to(engine.dev, non_blocking=True)
        res[k] = engine.pack_rgb8(img.reshape(-1, img.shape[-1]), (0.0, 1.0), bgr=bgr).reshape(H, W, -1).cpu().numpy()
    return res


# ---------------------------------------------------------------------------------------------------------------------
# consumer side: SaverMixin.save_image_grid / save_rgb_image / save_image (utils/mixins.py:43-58, 116-164) as used by
# test_step (systems/intrinsic_avatar.py:721-868)

# cv2.COLORMAP_JET as RGB triples (what ends up in the file when the reference passes applyColorMap's BGR image to
# cv2.imwrite); extracted with cv2.applyColorMap(np.arange(256)), zlib + base64
_JET_RGB = np.frombuffer(zlib.decompress(base64.b64decode(
    "eNod0gFHnAEAgOF3M8lkJklmkkkmSTJJkkmSySSZSTJJJpkkmWSSTJKcJGeSM8lJ5iRzkpxkTpIkySRnMidJkknSu+17fsMDwzAKYzABIZiGMMxCBOYhCksQgxWIwxokYBOSsA27sA+HcAQpOIE0nMEFXME13II8kEzJkseSI3nyVAqkUJ5LiZTJC6mUankpddIgjdIkLfJW2uSddMp76ZFe6ZePMiTDMipjMiEhmZawzEpE5iUqSxKTFYnLmiRkU5KyLbuyL4dyJCk5kbScyYVcybXcet+7DG8e+ueRl9me53r6xN/5/nrmcZE/iz0oda/cnQq3qvxR40at6/WuvvL7a5eb/fbGxVYX2v3a4VyXX7qd+eBUn5MDjg/6+ZMjI3767OC4A5P2Tflhxu4vds3Z8dX2BVsXffPN5mVff/fVqvXr1m5Y88OqLSt2LN+z9MDinxYd++yX+b99cmruudmXPvrjwxsz7rznLV7jFV7gGabxBFN4hIe4j7u4jUncxASuYRxXMIZLGMV5jOAshnEaQziBYziKwziEH7Efe7EH32MnvsM2fIst2ISN2IB1+BKrsRJfYBmW4HMsxAJ8inmYg48xCzPxwb86/wNdB5kugljpIFkqCHcY5NsNIiaDlIkgaDzIGgviRoPEkSB0OMgdCqKPBen/1f8LeOptzw==")),
    np.uint8).reshape(256, 3)


def png_bytes(img: np.ndarray, level: int = 1) -> bytes:
    """uint8 [H, W, 3 | 4] (RGB / RGBA, the order of the FILE) -> PNG.  cv2 (what the reference writes with) when it is
    importable, otherwise a zlib-only encoder (filter 0 on every row)."""
    img = np.ascontiguousarray(img)
    assert img.dtype == np.uint8 and img.ndim == 3 and img.shape[2] in (3, 4)
    try:
        import cv2
        code = cv2.COLOR_RGB2BGR if img.shape[2] == 3 else cv2.COLOR_RGBA2BGRA
        ok, buf = cv2.imencode(".png", cv2.cvtColor(img, code), [cv2.IMWRITE_PNG_COMPRESSION, level])
        if ok:
            return buf.tobytes()
    except ImportError:
        pass
    H, W, C = img.shape
    raw = np.concatenate([np.zeros((H, 1), np.uint8), img.reshape(H, W * C)], axis=1).tobytes()

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xffffffff)
    ihdr = struct.pack(">IIBBBBB", W, H, 8, 2 if C == 3 else 6, 0, 0, 0)
    return b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", ihdr) + chunk(b"IDAT", zlib.compress(raw, level)) + chunk(b"IEND", b"")


def exr_bytes(img: np.ndarray) -> bytes:
    """float32 [H, W, 3] (RGB) -> OpenEXR 2.0 single-part scanline file, 32-bit float channels B, G, R, no compression
    (what pyexr.write(path, img) stores for the 'hdr' grid of test_epoch_end, systems/intrinsic_avatar.py:875-879)."""
    img = np.ascontiguousarray(img, np.float32)
    H, W, C = img.shape
    assert C == 3

    def attr(name, typ, data):
        return name.encode() + b"\0" + typ.encode() + b"\0" + struct.pack("<i", len(data)) + data
    chlist = b"".join(c + b"\0" + struct.pack("<iBBBBii", 2, 0, 0, 0, 0, 1, 1) for c in (b"B", b"G", b"R")) + b"\0"
    box = struct.pack("<iiii", 0, 0, W - 1, H - 1)
    head = (struct.pack("<II", 20000630, 2) + attr("channels", "chlist", chlist) + attr("compression", "compression", b"\0")
            + attr("dataWindow", "box2i", box) + attr("displayWindow", "box2i", box) + attr("lineOrder", "lineOrder", b"\0")
            + attr("pixelAspectRatio", "float", struct.pack("<f", 1.0)) + attr("screenWindowCenter", "v2f", struct.pack("<ff", 0, 0))
            + attr("screenWindowWidth", "float", struct.pack("<f", 1.0)) + b"\0")
    line = 8 + 3 * W * 4
    first = len(head) + 8 * H
    table = struct.pack("<%dQ" % H, *[first + y * line for y in range(H)])
    bgr = img[:, :, ::-1].transpose(0, 2, 1)          # [H, channel (B, G, R), W]
    rows = np.empty((H, line), np.uint8)
    rows[:, :8] = np.frombuffer(b"".join(struct.pack("<ii", y, 3 * W * 4) for y in range(H)), np.uint8).reshape(H, 8)
    rows[:, 8:] = np.ascontiguousarray(bgr).view(np.uint8).reshape(H, 3 * W * 4)
    return head + table + rows.tobytes()


class FrameWriter:
    """The image consumer of a test / predict loop, off the render thread.

    The reference converts every output image to uint8 on the host (float32 D2H of every buffer, numpy clip / scale,
    cv2.cvtColor), concatenates the columns and calls cv2.imwrite inside test_step, i.e. between two frames.  Here
    ``save_image_grid`` packs every column straight into ONE uint8 grid on the device (``ia_pack_grid8``), starts an
    asynchronous copy into a pinned staging buffer on a side stream and returns; a worker thread waits for the copy's
    event, encodes the PNG (cv2 / zlib release the GIL) and writes it, together with the per-column images and the
    RGBA variants test_step saves from the same columns (systems/intrinsic_avatar.py:845-866).  The next frame's
    prepare + forward are enqueued while that happens; ``depth`` staging buffers bound the frames in flight.
    """

    def __init__(self, engine, save_dir: str, workers: int = 2, depth: int = 3, png_level: int = 1):
        self.engine, self.save_dir, self.png_level = engine, save_dir, png_level
        self._copy_stream = torch.cuda.Stream(device=engine.dev)
        self._free = queue.Queue()
        for _ in range(depth):
            self._free.put({})                 # staging slots: pinned host buffers by size, allocated on first use
        self._jobs = queue.Queue()
        self._errors = []
        self._lut_jet = torch.from_numpy(_JET_RGB.copy()).to(engine.dev)
        self._threads = [threading.Thread(target=self._work, daemon=True) for _ in range(workers)]
        for t in self._threads:
            t.start()

    # ---- paths
    def get_save_path(self, filename):
        path = os.path.join(self.save_dir, filename)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        return path

    # ---- device side
    def _grid(self, imgs):
        """[(kind, img [H, W, C], data_range, lut)] per 3-wide column, as get_image_grid_ lays them out."""
        cols = []
        for col in imgs:
            kw = dict(col.get("kwargs", {}))
            img = col["img"]
            if col["type"] == "rgb":
                if kw.get("data_format", "CHW") == "CHW":
                    img = img.permute(1, 2, 0)
                rng = kw.get("data_range", (0, 1))
                for s0 in range(0, img.shape[-1], 3):           # more than 3 channels: one column per 3 (utils/mixins.py:50)
                    cols.append(("rgb", img[..., s0:s0 + 3], rng, None))
            elif col["type"] == "grayscale":
                cmap = kw.get("cmap", "jet")
                if cmap not in (None, "jet"):
                    raise NotImplementedError(f"grayscale cmap {cmap!r}")
                cols.append(("grayscale", img, kw.get("data_range", None), self._lut_jet if cmap == "jet" else None))
            else:
                raise NotImplementedError(f"image grid column type {col['type']!r} (uint8 grids hold rgb / grayscale)")
        H = int(cols[0][1].shape[0])
        widths = [int(c[1].shape[1]) for c in cols]
        grid = torch.empty(H, sum(widths), 3, dtype=torch.uint8, device=self.engine.dev)
        x0 = 0
        for (kind, img, rng, lut), w in zip(cols, widths):
            self.engine.pack_grid8(grid, x0, img, kind=kind, data_range=rng, lut=lut)
            x0 += w
        return grid, widths

    def _stage(self, dev_tensors):
        """Async D2H of uint8 device tensors into a free staging slot; returns (slot, host views, event)."""
        slot = self._free.get()                    # blocks when `depth` frames are still being written
        done = torch.cuda.Event()
        cur = torch.cuda.current_stream(self.engine.dev)
        self._copy_stream.wait_stream(cur)
        host = []
        with torch.cuda.stream(self._copy_stream):
            for i, t in enumerate(dev_tensors):
                key = (i, tuple(t.shape))
                if key not in slot:
                    slot[key] = torch.empty(t.shape, dtype=torch.uint8, pin_memory=True)
                slot[key].copy_(t, non_blocking=True)
                t.record_stream(self._copy_stream)
                host.append(slot[key])
            done.record(self._copy_stream)
        return slot, host, done

    # ---- the reference's calls
    def save_image_grid(self, filename, imgs, captions=None, column_pattern=None, alpha=None, alpha_pattern=None):
        """SaverMixin.save_image_grid(filename, imgs) for uint8 grids.  With ``captions`` (one per column) the columns
        are also written one by one to ``column_pattern.format(caption=...)`` and, with ``alpha`` [H, W] (opacity), as
        RGBA to ``alpha_pattern.format(caption=...)`` -- the two loops that follow the grid in test_step."""
        grid, widths = self._grid(imgs)
        dev = [grid]
        if alpha is not None:
            a8 = torch.empty(grid.shape[0], int(alpha.shape[-1]), 3, dtype=torch.uint8, device=self.engine.dev)
            self.engine.pack_grid8(a8, 0, alpha.reshape(grid.shape[0], -1), kind="grayscale", data_range=(0, 1))
            dev.append(a8)
        slot, host, done = self._stage(dev)
        self._jobs.put(("grid", slot, host, done, self.get_save_path(filename), widths, captions, column_pattern, alpha_pattern))

    def save_rgb_image(self, filename, img, data_format="CHW", data_range=(0, 1)):
        self.save_image_grid(filename, [{"type": "rgb", "img": img, "kwargs": {"data_format": data_format, "data_range": data_range}}])

    def save_exr(self, filename, img):
        """The 'hdr' grid (float32 [H, W, 3]): written as 32-bit float OpenEXR by the worker."""
        host = img.detach().to("cpu", torch.float32).numpy()
        self._jobs.put(("exr", None, host, None, self.get_save_path(filename), None, None, None, None))

    # ---- host side
    def _work(self):
        while True:
            job = self._jobs.get()
            if job is None:
                return
            kind, slot, host, done, path, widths, captions, col_pat, alpha_pat = job
            try:
                if kind == "exr":
                    with open(path, "wb") as f:
                        f.write(exr_bytes(host))
                    continue
                done.synchronize()
                grid = host[0].numpy()
                with open(path, "wb") as f:
                    f.write(png_bytes(grid, self.png_level))
                if captions:
                    x0 = 0
                    for w, cap in zip(widths, captions):
                        col = grid[:, x0:x0 + w]
                        x0 += w
                        if col_pat:
                            with open(self.get_save_path(col_pat.format(caption=cap)), "wb") as f:
                                f.write(png_bytes(col, self.png_level))
                        if alpha_pat and len(host) > 1:
                            rgba = np.concatenate([col, host[1].numpy()[:, :, :1]], axis=-1)
                            with open(self.get_save_path(alpha_pat.format(caption=cap)), "wb") as f:
                                f.write(png_bytes(rgba, self.png_level))
            except Exception as e:            # surfaced by flush()
                self._errors.append(e)
            finally:
                if slot is not None:
                    self._free.put(slot)
                self._jobs.task_done()

    def flush(self):
        self._jobs.join()
        if self._errors:
            raise self._errors.pop(0)

    def close(self):
        self.flush()
        for _ in self._threads:
            self._jobs.put(None)
        for t in self._threads:
            t.join()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def render_sequence(model, frames, writer: FrameWriter, step: int = 0, H: int | None = None, W: int | None = None,
                    relight: bool = True):
    """The reference's test loop over an animation (Lightning calls test_step per frame, systems/intrinsic_avatar.py:
    586-868: prepare, forward, metrics, save_image_grid, per-column saves) with the consumer off the critical path:
    frame i's grid is packed on the device and handed to the writer's side stream / worker threads, then frame i + 1's
    prepare + forward are issued -- the D2H copy, the PNG encoding and the file writes of frame i overlap the kernels
    of frame i + 1.  Yields (index, device output dict) per frame; the writer is flushed at the end."""
    H = frames.H if H is None else H
    W = frames.W if W is None else W
    caps = ["rf"] + (["pbr", "pbr_demod", "albedo", "roughness", "metallic"] if relight else []) + ["depth", "normal"]
    for batch in frames:
        idx = int(batch["index"])
        model.prepare(batch if relight else {k: v for k, v in batch.items() if k != "hdri"})
        out = model.forward(batch["rays"], move_to_cpu=False)
        v = lambda k, c: out[k].reshape(H, W, c)
        cols = [{"type": "rgb", "img": v("comp_rgb_full", 3), "kwargs": {"data_format": "HWC"}}]
        if relight:
            cols += [{"type": "rgb", "img": v("comp_rgb_phys_full", 3), "kwargs": {"data_format": "HWC"}},
                     {"type": "rgb", "img": v("comp_demod_phys_full", 3), "kwargs": {"data_format": "HWC"}},
                     {"type": "rgb", "img": v("comp_albedo_full", 3), "kwargs": {"data_format": "HWC"}},
                     {"type": "grayscale", "img": v("comp_roughness_full", 1)[..., 0], "kwargs": {"data_range": (0, 1), "cmap": None}},
                     {"type": "grayscale", "img": v("comp_metallic_full", 1)[..., 0], "kwargs": {"data_range": (0, 1), "cmap": None}}]
        cols += [{"type": "grayscale", "img": v("depth", 1)[..., 0], "kwargs": {}},
                 {"type": "rgb", "img": v("comp_normal", 3), "kwargs": {"data_format": "HWC", "data_range": (-1, 1)}}]
        writer.save_image_grid(f"it{step}-test-all/{idx}.png", cols, captions=caps,
                               column_pattern=f"it{step}-test/{idx:04}-{{caption}}.png",
                               alpha=out["opacity"].reshape(H, W).clamp(0, 1),
                               alpha_pattern=f"it{step}-test-with-alpha/{idx:04}-{{caption}}.png")
        yield idx, out
    writer.flush()
