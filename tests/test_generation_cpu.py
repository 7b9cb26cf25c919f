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

    def llm_expand(self, src_idx, new_B):
        self.cache.reorder(src_idx.long())
        self.mask = self.mask.index_select(0, src_idx.long())

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


    def beam_topk(self, logits, running_scores, batch, beams, keep, ban_id=-1, prev_tokens=None, prev_len=0,
                  repetition_penalty=1.0, penalty_on_logits=False, no_repeat_ngram=0, allowed=None):
        return torch_beam_topk(logits, running_scores, batch, beams, keep, ban_id, prev_tokens, prev_len, repetition_penalty,
                               penalty_on_logits, no_repeat_ngram, allowed)

    def beam_state(self, batch, beams, max_length, pad_token_id, device):
        return TorchBeamState(batch, beams, max_length, pad_token_id)

    def beam_step(self, st, topk_lp, topk_idx, cur_len, eos_token_id, length_penalty, early_stopping):
        st.step(topk_lp, topk_idx, self.cfg.llm_vocab, cur_len, eos_token_id, length_penalty, early_stopping)

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


def torch_beam_topk(logits, running_scores, batch, beams, keep, ban_id=-1, prev_tokens=None, prev_len=0,
                    repetition_penalty=1.0, penalty_on_logits=False, no_repeat_ngram=0, allowed=None):
    """torch formulation of the HF `_beam_search` step that emu_beam_topk implements on the device (HF processor classes
    restated: RepetitionPenalty, NoRepeatNGram, MinLength, PrefixConstrained)."""
    V = logits.shape[-1]
    x = logits.float().clone()
    prev = prev_tokens[:, :prev_len].long() if (prev_tokens is not None and prev_len > 0) else None

    def rep(t):
        sc = torch.gather(t, 1, prev)
        sc = torch.where(sc < 0, sc * repetition_penalty, sc / repetition_penalty)
        return t.scatter(1, prev, sc)
    if prev is not None and repetition_penalty != 1.0 and penalty_on_logits:
        x = rep(x)
    lp = torch.log_softmax(x, dim=-1)
    if prev is not None and repetition_penalty != 1.0 and not penalty_on_logits:
        lp = rep(lp)
    if prev is not None and no_repeat_ngram and prev_len + 1 >= no_repeat_ngram:
        n = no_repeat_ngram
        for r in range(prev.shape[0]):
            row = prev[r].tolist()
            tail = row[len(row) - (n - 1):] if n > 1 else []
            for i in range(len(row) - n + 1):
                if row[i:i + n - 1] == tail:
                    lp[r, row[i + n - 1]] = float("-inf")
    if ban_id is not None and ban_id >= 0:
        lp[:, ban_id] = float("-inf")
    if allowed is not None:
        lp = lp.masked_fill(allowed.to(lp.device) == 0, float("-inf"))
    lp = lp.view(batch, beams, V) + running_scores.view(batch, beams)[:, :, None]
    v, i = torch.topk(lp.view(batch, beams * V), k=keep)
    return v, i.to(torch.int32)


def _gather_beams(t, idx):
    while idx.dim() < t.dim():
        idx = idx.unsqueeze(-1)
    return torch.take_along_dim(t, idx, dim=1)


class TorchBeamState:
    """torch formulation of the hypothesis bookkeeping of HF's vectorised `_beam_search` (transformers >= 4.50) — what
    emu_beam_step implements on the device.  Same attribute surface as emu_b200._lib.BeamState; pinned against the
    reference's own lm.generate by the golden-id tests below, and the device kernel is pinned against THIS
    (tests/test_ops_gpu.py::test_beam_step_matches_torch_formulation)."""

    def __init__(self, batch, beams, max_length, pad):
        self.batch, self.beams, self.max_length = batch, beams, max_length
        self.running_seq = torch.full((2, batch, beams, max_length), pad, dtype=torch.int32)
        self.sequences = torch.full((2, batch, beams, max_length), pad, dtype=torch.int32)
        self.running_scores = torch.zeros(batch, beams)
        self.running_scores[:, 1:] = -1e9
        self.beam_scores = torch.full((batch, beams), -1e9)
        self.is_finished = torch.zeros(batch, beams, dtype=torch.int32)
        self.fin_len = torch.zeros(batch, beams, dtype=torch.int32)
        self.unsat = torch.ones(batch, dtype=torch.int32)
        self.done = torch.zeros(1, dtype=torch.int32)
        self.next_tokens = torch.zeros(batch * beams, dtype=torch.int32)
        self.beam_src = torch.zeros(batch * beams, dtype=torch.int32)

    def live(self, cur_len):
        return cur_len & 1

    def is_done(self):
        return bool(self.done.item())

    def result(self, cur_len):
        best = self.sequences[self.live(cur_len), :, 0, :]
        return best[:, :int(self.fin_len[:, 0].max())].to(torch.int64)

    def step(self, topk_lp, topk_i, V, cur_len, eos, length_penalty, early_stopping):
        if self.is_done():
            return
        nb, L = self.beams, self.max_length
        pin, pout = cur_len & 1, (cur_len + 1) & 1
        running_seq, sequences = self.running_seq[pin].long(), self.sequences[pin].long()
        is_finished, unsat = self.is_finished.bool(), self.unsat.bool()[:, None]
        top_mask = torch.cat((torch.ones(nb, dtype=torch.bool), torch.zeros(nb, dtype=torch.bool)))
        topk_i = topk_i.long()
        topk_beam, topk_ids = topk_i // V, topk_i % V
        topk_seq = _gather_beams(running_seq, topk_beam)
        topk_seq[:, :, cur_len] = topk_ids
        hits = (topk_ids == eos) | (cur_len + 1 >= L)
        run_lp = topk_lp + hits.float() * -1.0e9
        nxt_i = torch.topk(run_lp, k=nb)[1]
        self.running_seq[pout] = _gather_beams(topk_seq, nxt_i).to(torch.int32)
        self.running_scores = _gather_beams(run_lp, nxt_i)
        just_fin = hits & top_mask[None, :]
        fin_lp = topk_lp / ((cur_len + 1) ** length_penalty)
        full = torch.all(is_finished, dim=-1, keepdim=True) & (early_stopping is True)
        fin_lp = fin_lp + full.float() * -1.0e9
        fin_lp = fin_lp + (~unsat).float() * -1.0e9
        fin_lp = fin_lp + (~just_fin).float() * -1.0e9
        m_seq = torch.cat((sequences, topk_seq), dim=1)
        m_sc = torch.cat((self.beam_scores, fin_lp), dim=1)
        m_fin = torch.cat((is_finished, just_fin), dim=1)
        m_len = torch.cat((self.fin_len.long(), torch.full_like(topk_ids, cur_len + 1)), dim=1)
        sel = torch.topk(m_sc, k=nb)[1]
        self.sequences[pout] = _gather_beams(m_seq, sel).to(torch.int32)
        self.beam_scores = _gather_beams(m_sc, sel)
        is_finished = _gather_beams(m_fin, sel)
        self.is_finished = is_finished.to(torch.int32)
        self.fin_len = _gather_beams(m_len, sel).to(torch.int32)
        batch_off = (torch.arange(self.batch) * nb).view(-1, 1)
        self.next_tokens = _gather_beams(topk_ids, nxt_i).reshape(-1).to(torch.int32)
        self.beam_src = (_gather_beams(topk_beam, nxt_i) + batch_off).reshape(-1).to(torch.int32)
        best_len = L if (early_stopping == "never" and length_penalty > 0.0) else cur_len + 1
        best_running = self.running_scores[:, :1] / (best_len ** length_penalty)
        worst_fin = torch.where(is_finished, self.beam_scores.min(dim=1, keepdim=True)[0], -1.0e9)
        unsat = unsat & torch.any(best_running > worst_fin, dim=-1, keepdim=True)
        self.unsat = unsat[:, 0].to(torch.int32)
        improvement = torch.any(unsat)
        open_beam = ~(torch.all(is_finished) & (early_stopping is True))
        valid = ~torch.all(hits)
        if not bool(improvement & open_beam & valid):
            self.done[0] = 1


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
