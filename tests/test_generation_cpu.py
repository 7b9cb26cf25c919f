"""CPU: the host-side generation control (emu_b200/generation.py: greedy / beam search restated from HF) driven by a
fake engine whose prefill/decode are the CPU oracle — token ids must equal what the reference's own
lm.generate produced (golden fixture)."""
import os
from types import SimpleNamespace

import pytest
import torch
import torch.nn.functional as F

from helpers import TINY_LLAMA, VOCAB, make_emu2_state_dict
from oracle import emu_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "emu2_tiny.pt")
L, NH = TINY_LLAMA["num_hidden_layers"], TINY_LLAMA["num_attention_heads"]


class OracleEngine:
    """Same method surface as emu_b200._lib.Engine, arithmetic by the oracle (test double, CPU)."""

    def __init__(self, sd, max_batch=8):
        self.sd = sd
        self.cfg = SimpleNamespace(llm_vocab=VOCAB, llm_max_batch=max_batch)
        self.cache = None
        self.mask = None

    def llm_reset(self):
        self.cache = None

    def llm_prefill(self, embeds, attention_mask, hf_positions=True, want_hidden=False, want_logits=True):
        self.cache = O.KVCache(L)
        self.mask = attention_mask.clone()
        pos = O.hf_position_ids(self.mask) if hf_positions else None
        h = O.llama_forward(self.sd, embeds, self.mask, layers=L, heads=NH, position_ids=pos, cache=self.cache)
        self.hf = hf_positions
        return (h if want_hidden else None), O.lm_logits(self.sd, h[:, -1]).float()

    def llm_decode(self, token_ids=None, embeds=None, beam_src=None, logits=None, hidden=None, next_ids=None,
                   ban_id=-1, B=None):
        if beam_src is not None:
            self.cache.reorder(beam_src.long())
            self.mask = self.mask.index_select(0, beam_src.long())
        e = F.embedding(token_ids.long(), self.sd["decoder.lm.model.embed_tokens.weight"]).unsqueeze(1) \
            if token_ids is not None else embeds.unsqueeze(1)
        self.mask = torch.cat((self.mask, torch.ones(self.mask.shape[0], 1, dtype=self.mask.dtype)), dim=1)
        pos = (self.mask.long().sum(-1, keepdim=True) - 1) if self.hf else None
        h = O.llama_forward(self.sd, e, self.mask, layers=L, heads=NH, position_ids=pos, cache=self.cache)
        lg = O.lm_logits(self.sd, h[:, -1]).float()
        if logits is not None:
            logits.copy_(lg)
        if hidden is not None:
            hidden.copy_(h[:, -1])
        if next_ids is not None:
            if ban_id >= 0:
                lg[:, ban_id] = float("-inf")
            next_ids.copy_(lg.argmax(-1).to(next_ids.dtype))


    def beam_topk(self, logits, running_scores, batch, beams, keep, ban_id=-1, prev_tokens=None, repetition_penalty=1.0):
        """torch formulation of the HF _beam_search step that emu_beam_topk implements on the device"""
        V = logits.shape[-1]
        lp = torch.log_softmax(logits.float(), dim=-1)
        if prev_tokens is not None and repetition_penalty != 1.0:
            sc = torch.gather(lp, 1, prev_tokens)
            sc = torch.where(sc < 0, sc * repetition_penalty, sc / repetition_penalty)
            lp = lp.scatter(1, prev_tokens, sc)
        if ban_id is not None and ban_id >= 0:
            lp[:, ban_id] = float("-inf")
        lp = lp.view(batch, beams, V) + running_scores[:, :, None]
        return torch.topk(lp.view(batch, beams * V), k=keep)

    def sample_tokens(self, logits, temperature=1.0, top_k=0, top_p=1.0, ban_id=-1, seed=0, offset=0):
        """HF warpers + torch.multinomial (what emu_sample_tokens implements on the device)"""
        scores = logits.float().clone()
        if ban_id is not None and ban_id >= 0:
            scores[:, ban_id] = float("-inf")
        scores = scores / temperature
        if top_k:
            kth = torch.topk(scores, min(top_k, scores.shape[-1]))[0][..., -1, None]
            scores = scores.masked_fill(scores < kth, float("-inf"))
        if top_p < 1.0:
            s_sorted, s_idx = torch.sort(scores, descending=False)
            remove = s_sorted.softmax(-1).cumsum(-1) <= (1 - top_p)
            remove[..., -1:] = False
            scores = scores.masked_fill(remove.scatter(1, s_idx, remove), float("-inf"))
        g = torch.Generator().manual_seed((seed + offset) % (2 ** 63))
        return torch.multinomial(scores.softmax(-1), 1, generator=g).squeeze(1).to(torch.int32)


@pytest.fixture(scope="module")
def setup():
    gold = torch.load(GOLD)
    sd = make_emu2_state_dict()
    e = O.encode_image(sd, gold["image"], patch=14, num_heads=4, layers=2, n_query=4)
    pie = F.linear(e.view(-1, e.shape[-1]), sd["project_up.weight"])
    emb = O.splice_embeds(sd, gold["gen_input_ids"], pie, 32003)
    return gold, sd, emb


def test_greedy_matches_reference_generate(setup):
    from emu_b200 import generation
    gold, sd, emb = setup
    toks = generation.greedy_search(OracleEngine(sd), emb, gold["gen_attention_mask"], 12, 2, 32000, min_length=1)
    assert torch.equal(toks, gold["gen_ids_greedy"])


def test_beam5_default_length_penalty_matches_reference_generate(setup):
    from emu_b200 import generation
    gold, sd, emb = setup
    toks = generation.beam_search(OracleEngine(sd), emb[:1], gold["gen_attention_mask"][:1], 5, 12, 2, 32000,
                                  min_length=1, length_penalty=-1)
    assert torch.equal(toks, gold["gen_ids_beam5"])


def test_beam3_batch2_matches_reference_generate(setup):
    from emu_b200 import generation
    gold, sd, emb = setup
    toks = generation.beam_search(OracleEngine(sd), emb, gold["gen_attention_mask"], 3, 12, 2, 32000, min_length=1,
                                  length_penalty=1.0)
    assert torch.equal(toks, gold["gen_ids_beam3_lp1"])


def test_beam_batch_limit(setup):
    from emu_b200 import generation
    gold, sd, emb = setup
    with pytest.raises(ValueError):
        generation.beam_search(OracleEngine(sd, max_batch=4), emb, gold["gen_attention_mask"], 5, 4, 2, 32000)
