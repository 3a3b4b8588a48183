"""Offline dataset darkener on the GPU -- host-side mirror of
``apply_lowlight_and_save(input_dir, output_dir, lowlight_param=7.5, batch_size=16)`` (utils/lowlight_process.py:10-87).

    JPEG bytes --nvJPEG decode (GPU)--> uint8 RGB [3,H,W] --dd_synth_fwd (truncating uint8 writer)--> darkened uint8
               --nvJPEG encode (GPU, quality 95 = cv2.imwrite's default)--> output_dir/<same file name>

The arithmetic between the codecs is the library's own kernel (``apply_lowlight``: ``(pow(u8/255, p) * 255).astype(uint8)``,
utils/lowlight_process.py:68,74), bit-exact against the reference on the same decoded pixels.  The codec is nvJPEG, reached
through ``torchvision.io.decode_jpeg / encode_jpeg(device='cuda')`` (a library, like the reference's PIL / OpenCV codecs):
decoders differ by a level or two on some pixels and encoders in their Huffman/quantisation choices, so files are not
byte-identical to ``cv2.imwrite``'s -- parity is stated on the pre-encode uint8 arrays (tests/test_offline.py).  PNG inputs
(the reference also accepts them) are decoded on the host and join the same GPU path.  There is no CPU fallback for the
darkening itself.
"""
from __future__ import annotations

import os
from typing import Iterable, List, Optional

import torch

from .lowlight import apply_lowlight

IMG_EXTENSIONS = (".jpg", ".jpeg", ".png")   # utils/lowlight_process.py:24


def _read(path: str) -> torch.Tensor:
    with open(path, "rb") as f:
        return torch.frombuffer(bytearray(f.read()), dtype=torch.uint8)


def decode_rgb(paths: Iterable[str], device) -> List[torch.Tensor]:
    """Files -> list of uint8 RGB [3,H,W] CUDA tensors (JPEG: nvJPEG on ``device``; PNG: host decode, then copied)."""
    from torchvision.io import ImageReadMode, decode_image, decode_jpeg

    paths = list(paths)
    out: List[Optional[torch.Tensor]] = [None] * len(paths)
    jpeg_idx = [i for i, q in enumerate(paths) if q.lower().endswith((".jpg", ".jpeg"))]
    if jpeg_idx:
        datas = [_read(paths[i]) for i in jpeg_idx]
        imgs = decode_jpeg(datas, mode=ImageReadMode.RGB, device=device)   # batched nvJPEG decode
        for i, im in zip(jpeg_idx, imgs):
            out[i] = im
    for i, q in enumerate(paths):
        if out[i] is None:
            out[i] = decode_image(_read(q), mode=ImageReadMode.RGB).to(device)
    return out  # type: ignore[return-value]


def darken_images(images: List[torch.Tensor], lowlight_param: float = 7.5, lut: Optional[torch.Tensor] = None) -> List[torch.Tensor]:
    """uint8 RGB [3,H,W] CUDA tensors (any mix of sizes) -> darkened uint8 tensors: one ``dd_synth_fwd`` launch over the
    concatenated bytes (the pass is elementwise, so images of different resolutions share it -- the reference needs one batch
    per resolution group, utils/lowlight_process.py:34-40)."""
    if not images:
        return []
    flat = torch.cat([im.reshape(-1) for im in images])
    pad = (-flat.numel()) % 16                         # the kernel wants 16-byte aligned buffers; any length
    if pad:
        flat = torch.cat([flat, flat.new_zeros(pad)])
    dark = apply_lowlight(flat.view(1, 1, 1, -1), lowlight_param, lut=lut).view(-1)
    outs, off = [], 0
    for im in images:
        n = im.numel()
        outs.append(dark[off:off + n].view(im.shape))
        off += n
    return outs


def darken_directory(input_dir: str, output_dir: str, lowlight_param: float = 7.5, batch_size: int = 16, device="cuda:0",
                     quality: int = 95, verbose: bool = False) -> int:
    """Same contract as the reference's ``apply_lowlight_and_save``: every ``.jpg/.jpeg/.png`` of ``input_dir`` is darkened at
    its own resolution and written under the same file name to ``output_dir``.  Returns the number of images written.
    Outputs are always JPEG-encoded for ``.jpg/.jpeg`` names and PNG-encoded (host) for ``.png`` names, as cv2.imwrite does."""
    from torchvision.io import encode_jpeg, encode_png

    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("darken_directory needs a CUDA device (no CPU fallback)")
    os.makedirs(output_dir, exist_ok=True)
    names = sorted(f for f in os.listdir(input_dir) if f.lower().endswith(IMG_EXTENSIONS))
    if not names:
        if verbose:
            print("no images found in the input directory")
        return 0
    done = 0
    for b0 in range(0, len(names), batch_size):
        chunk = names[b0:b0 + batch_size]
        imgs = decode_rgb([os.path.join(input_dir, n) for n in chunk], dev)
        dark = darken_images(imgs, lowlight_param)
        jpeg = [i for i, n in enumerate(chunk) if not n.lower().endswith(".png")]
        if jpeg:
            enc = encode_jpeg([dark[i].contiguous() for i in jpeg], quality=quality)   # CUDA tensors in -> nvJPEG encode
            for i, data in zip(jpeg, enc):
                with open(os.path.join(output_dir, chunk[i]), "wb") as f:
                    f.write(data.cpu().numpy().tobytes())
        for i, n in enumerate(chunk):
            if n.lower().endswith(".png"):
                with open(os.path.join(output_dir, n), "wb") as f:
                    f.write(encode_png(dark[i].cpu()).numpy().tobytes())
        done += len(chunk)
        if verbose:
            print(f"darkened {done}/{len(names)} images")
    return done
