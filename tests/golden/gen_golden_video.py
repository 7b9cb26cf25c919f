"""Generates tests/golden/emu2_tiny_video.pt: `EmuModel.generate` of the UNMODIFIED reference with BOTH an image and a video in
the prompt (Emu2/emu/emu.py:197-211: the video frames go through encode_image with v_query tokens per frame and fill the
[gIMG] slots, the picture fills the <image> slots), fp32 on CPU, greedy and 3-beam; the new token ids are captured at the
tokenizer's batch_decode.  Same tiny model / weights as gen_golden.py.

Run in the authoring container only:  python tests/golden/gen_golden_video.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import TINY_LLAMA, TINY_VISION, make_emu2_state_dict  # noqa: E402
from oracle import ref_shim  # noqa: E402


def main():
    torch.manual_seed(0)
    d = ref_shim.make_llama_config_dir(TINY_LLAMA["hidden_size"], TINY_LLAMA["num_hidden_layers"],
                                       TINY_LLAMA["num_attention_heads"], TINY_LLAMA["intermediate_size"],
                                       max_pos=TINY_LLAMA["max_position_embeddings"])
    vk = dict(TINY_VISION)
    vk.pop("patch_size")
    model = ref_shim.build_emu2_model(dict(vk, patch_size=14), d)
    model.load_state_dict(make_emu2_state_dict(), strict=True)
    tok = model.decoder.tokenizer
    captured = []
    decode = tok.batch_decode
    tok.batch_decode = lambda ids, **kw: (captured.append(ids.clone()), decode(ids, **kw))[1]
    g = torch.Generator().manual_seed(99)
    image = torch.randn(1, 3, 56, 56, generator=g)
    video = torch.randn(3, 3, 56, 56, generator=g)                      # three frames
    texts = ["[<IMG_PLH>]A picture. [<VID_PLH>][<VID_PLH>][<VID_PLH>]What happens in the video?"]
    t2 = [t.replace("[<IMG_PLH>]", model.image_placeholder).replace("[<VID_PLH>]", model.video_placeholder) for t in texts]
    enc = tok(t2, padding="longest", return_tensors="pt")
    out = {"image": image, "video": video, "input_ids": enc.input_ids, "attention_mask": enc.attention_mask}
    with torch.no_grad():
        for name, kw in (("greedy", dict(num_beams=1)), ("beam3", dict(num_beams=3, length_penalty=1.0))):
            text = model.generate(text=texts, image=image, video=video, max_new_tokens=10, **kw)
            out["ids_" + name], out["text_" + name] = captured[-1], text
            print(name, captured[-1].tolist())
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu2_tiny_video.pt")
    torch.save(out, path)
    print("wrote", path, tuple(enc.input_ids.shape), int((enc.input_ids == 32003).sum()), int((enc.input_ids == 32004).sum()))


if __name__ == "__main__":
    main()
