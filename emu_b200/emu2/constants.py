"""Special-token strings and preprocessing constants of Emu2 (values as published in Emu2/emu/constants.py)."""
EVA_IMAGE_SIZE = 448
OPENAI_DATASET_MEAN = (0.48145466, 0.4578275, 0.40821073)
OPENAI_DATASET_STD = (0.26862954, 0.26130258, 0.27577711)

DEFAULT_PAD_TOKEN = "[PAD]"
DEFAULT_BOS_TOKEN = "<s>"
DEFAULT_EOS_TOKEN = "</s>"
DEFAULT_UNK_TOKEN = "<unk>"
DEFAULT_IMG_TOKEN = "[IMG]"
DEFAULT_IMG_END_TOKEN = "[/IMG]"
DEFAULT_IMAGE_TOKEN = "<image>"
DEFAULT_gIMG_TOKEN = "[gIMG]"
DEFAULT_gIMG_END_TOKEN = "[/gIMG]"
DEFAULT_EOC_TOKEN = "[EOC]"
DEFAULT_VIDEO_TOKEN = "[VIDEO]"
GRD_SYMBOL = "<grounding>"
BOP_SYMBOL = "<phrase>"
EOP_SYMBOL = "</phrase>"
BOO_SYMBOL = "<object>"
EOO_SYMBOL = "</object>"
DOM_SYMBOL = "</delimiter_of_multi_objects/>"
REC_SYMBOL = "<REC>"
USER_TOKEN = "[USER]"
ASSISTANT_TOKEN = "[ASSISTANT]"
DEFAULT_IMG_PLACEHOLDER = "[<IMG_PLH>]"
DEFAULT_VID_PLACEHOLDER = "[<VID_PLH>]"
FAKE_VIDEO_END_TOKEN = "[/VIDEO]"
GROUND_SYSTEM_MESSAGE = "You are a helpful assistant, dedicated to provide concise and efficient answers."
SYSTEM_MESSAGE = "You are a helpful assistant, dedicated to delivering comprehensive and meticulous responses."


def special_token_list(instruct=False, quantized_size=256):
    """Order matters: ids are assigned sequentially after the 32000 LLaMA pieces + [PAD] (Emu2/emu/lm.py:12-65)."""
    toks = [DEFAULT_IMG_TOKEN, DEFAULT_IMG_END_TOKEN, DEFAULT_IMAGE_TOKEN, DEFAULT_gIMG_TOKEN, DEFAULT_gIMG_END_TOKEN,
            DEFAULT_EOC_TOKEN, DEFAULT_VIDEO_TOKEN, GRD_SYMBOL, BOP_SYMBOL, EOP_SYMBOL, BOO_SYMBOL, EOO_SYMBOL,
            DOM_SYMBOL, REC_SYMBOL]
    toks += ["<patch_index_%s>" % str(i).zfill(4) for i in range(quantized_size + 1)]
    if instruct:
        toks += [USER_TOKEN, ASSISTANT_TOKEN]
    return toks
