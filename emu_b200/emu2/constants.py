"""Special-token strings and preprocessing constants of Emu2.

The VALUES are fixed by the published checkpoints and tokenizer (Emu2/emu/constants.py, Emu2/emu/lm.py:12-65): change one
and token ids or pixel statistics no longer match the weights.  They are grouped here by what the engine does with them.
"""
# ---- image pre-processing (emu_preprocess_image): EVA-CLIP input side and the OpenAI CLIP pixel statistics ----
EVA_IMAGE_SIZE = 448
OPENAI_DATASET_MEAN, OPENAI_DATASET_STD = (0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711)

# ---- tokenizer specials.  [PAD] is piece 32000; the rest get consecutive ids in the order of special_token_list() ----
DEFAULT_PAD_TOKEN, DEFAULT_BOS_TOKEN, DEFAULT_EOS_TOKEN, DEFAULT_UNK_TOKEN = "[PAD]", "<s>", "</s>", "<unk>"
# an image span in the prompt is  [IMG] <image> x n_query [/IMG]  (video frames use the [gIMG] pair), ids 32001..32005
DEFAULT_IMG_TOKEN, DEFAULT_IMAGE_TOKEN, DEFAULT_IMG_END_TOKEN = "[IMG]", "<image>", "[/IMG]"
DEFAULT_gIMG_TOKEN, DEFAULT_gIMG_END_TOKEN = "[gIMG]", "[/gIMG]"
DEFAULT_EOC_TOKEN, DEFAULT_VIDEO_TOKEN, FAKE_VIDEO_END_TOKEN = "[EOC]", "[VIDEO]", "[/VIDEO]"
# grounding vocabulary (phrase / object brackets, multi-object delimiter, referring-expression marker)
GRD_SYMBOL, REC_SYMBOL = "<grounding>", "<REC>"
BOP_SYMBOL, EOP_SYMBOL = "<phrase>", "</phrase>"
BOO_SYMBOL, EOO_SYMBOL, DOM_SYMBOL = "<object>", "</object>", "</delimiter_of_multi_objects/>"
# chat roles (only in the instruct tokenizer)
USER_TOKEN, ASSISTANT_TOKEN = "[USER]", "[ASSISTANT]"

# ---- what callers write in a prompt; EmuModel swaps them for the real spans before tokenising ----
DEFAULT_IMG_PLACEHOLDER, DEFAULT_VID_PLACEHOLDER = "[<IMG_PLH>]", "[<VID_PLH>]"

# ---- system prompts of the chat pipeline (plain / grounding) ----
SYSTEM_MESSAGE = "You are a helpful assistant, dedicated to delivering comprehensive and meticulous responses."
GROUND_SYSTEM_MESSAGE = "You are a helpful assistant, dedicated to provide concise and efficient answers."


def special_token_list(instruct=False, quantized_size=256):
    """Order matters: ids are assigned sequentially after the 32000 LLaMA pieces + [PAD] (Emu2/emu/lm.py:12-65)."""
    toks = [DEFAULT_IMG_TOKEN, DEFAULT_IMG_END_TOKEN, DEFAULT_IMAGE_TOKEN, DEFAULT_gIMG_TOKEN, DEFAULT_gIMG_END_TOKEN,
            DEFAULT_EOC_TOKEN, DEFAULT_VIDEO_TOKEN, GRD_SYMBOL, BOP_SYMBOL, EOP_SYMBOL, BOO_SYMBOL, EOO_SYMBOL,
            DOM_SYMBOL, REC_SYMBOL]
    toks += ["<patch_index_%s>" % str(i).zfill(4) for i in range(quantized_size + 1)]
    if instruct:
        toks += [USER_TOKEN, ASSISTANT_TOKEN]
    return toks
