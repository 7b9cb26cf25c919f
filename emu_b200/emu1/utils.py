"""Input preparation of the Emu1 example scripts (Emu1/utils.py): a PIL image -> the [1, 3, 224, 224] fp32 tensor the model takes,
and 8 evenly spaced frames of a video as 8 such tensors.  Host code, outside every timed region; kept numerically identical to
the reference (Pillow's default resize filter, normalisation in float64, one cast to fp32 at the end)."""
import numpy as np
import torch
from PIL import Image

from ..emu2.constants import OPENAI_DATASET_MEAN, OPENAI_DATASET_STD

IMAGE_SIDE = 224                      # Emu-14B.json: vision_cfg.image_size


def get_index(num_frames, num_segments):
    """frame numbers at the centres of `num_segments` equal spans of a `num_frames` clip (utils.py:7-14)"""
    span = float(num_frames - 1) / num_segments
    first = int(span / 2)
    return np.array([first + int(np.round(span * s)) for s in range(num_segments)])


def process_img(img_path=None, img=None, device=torch.device("cuda")):
    """utils.py:17-30.  `img.resize` uses Pillow's default filter (bicubic) and keeps uint8; the mean / std arithmetic runs in
    float64 and is rounded to fp32 once, as the reference does it."""
    assert img_path is not None or img is not None, "you should pass either path to an image or a PIL image object"
    picture = Image.open(img_path).convert("RGB") if img_path else img
    pixels = np.asarray(picture.resize((IMAGE_SIDE, IMAGE_SIDE)), dtype=np.float64) / 255.0
    pixels = (pixels - np.asarray(OPENAI_DATASET_MEAN)) / np.asarray(OPENAI_DATASET_STD)
    return torch.from_numpy(pixels).to(device).to(torch.float).permute(2, 0, 1).unsqueeze(0)


def process_video(video_path=None, num_segments=8, image_placeholder=None, device=torch.device("cuda")):
    """utils.py:33-45: -> (list of frame tensors, the text with one image placeholder per frame).  Needs `decord`."""
    try:
        from decord import VideoReader
    except ImportError as ex:  # the reference imports decord at module level; here only video inputs need it
        raise ImportError("process_video needs the `decord` package to read video files") from ex
    if image_placeholder is None:
        from .inference import image_placeholder
    reader = VideoReader(video_path)
    frames = [process_img(img=Image.fromarray(reader[int(i)].asnumpy()).convert("RGB"), device=device)
              for i in get_index(len(reader), num_segments)]
    return frames, image_placeholder * len(frames)
