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
                               penalty_on_logits, no_repeat_ngram, allowed, write_back=True)

    def beam_state(self, batch, beams, max_length, pad_token_id, device):
        return TorchBeamState(batch, beams, max_length, pad_token_id)

    def beam_step(self, st, topk_lp, topk_idx, cur_len, eos_token_id, length_penalty, early_stopping):
        st.step(topk_lp, topk_idx, self.cfg.llm_vocab, cur_len, eos_token_id, length_penalty, early_stopping)
        st.done_calls.add_(st.done)       # as emu_b200._lib.Engine.beam_step does after the device kernel

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
                    repetition_penalty=1.0, penalty_on_logits=False, no_repeat_ngram=0, allowed=None, write_back=False):
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
    if write_back:   # emu_beam_topk works in place: `logits` holds the processed, running-score-shifted log-probabilities
        logits.copy_(lp.view(batch * beams, V))
    v, i = torch.topk(lp.view(batch, beams * V), k=keep)
    return v, i.to(torch.int32)


def _gather_beams(t, idx):
    while idx.dim() < t.dim():
        idx = idx.unsqueeze(-1)
    return torch.take_along_dim(t, idx, dim=1)


from emu_b200._lib import BeamState as _BeamStateCls  # noqa: E402  (importing _lib does not load the library)


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
        self.done_calls = torch.zeros(1, dtype=torch.int32)
        self.next_tokens = torch.zeros(batch * beams, dtype=torch.int32)
        self.beam_src = torch.zeros(batch * beams, dtype=torch.int32)

    def live(self, cur_len):
        return cur_len & 1

    def is_done(self):
        return bool(self.done.item())

    final_len = _BeamStateCls.final_len        # the product's own plane arithmetic (host code, device independent)

    def result(self, cur_len, n=1):
        cur_len = self.final_len(cur_len)
        best = self.sequences[self.live(cur_len), :, :n, :].reshape(self.batch * n, self.max_length)
        return best[:, :int(self.fin_len[:, :n].max())].to(torch.int64)

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


# ---- generation control beyond plain greedy / beam search: pinned to the reference's own lm.generate (golden ids made by
# ---- tests/golden/gen_golden_control.py from the unmodified reference decoder on the same tiny model) -------------------
CTRL = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "emu2_tiny_control.pt")
SMALL = [6554, 18722, 15312, 29412, 3895, 1741]
ALLOWED = list(range(100, 140)) + [2]


def small_fn(batch_id, ids):
    return SMALL + ([2] if len(ids) >= 8 else [])


def prefix_fn(batch_id, ids):
    return [2] if len(ids) >= 3 else ALLOWED


@pytest.fixture(scope="module")
def ctrl():
    return torch.load(CTRL)


@pytest.mark.parametrize("name,kw", [
    ("greedy_rep", dict(repetition_penalty=1.5)),
    ("greedy_ngram2", dict(no_repeat_ngram_size=2)),
    ("greedy_rep_ngram", dict(repetition_penalty=1.3, no_repeat_ngram_size=3)),
    ("greedy_prefix", dict(prefix_allowed_tokens_fn=prefix_fn)),
    ("greedy_small", dict(prefix_allowed_tokens_fn=small_fn)),
    ("greedy_small_rep", dict(prefix_allowed_tokens_fn=small_fn, repetition_penalty=1.5)),
    ("greedy_small_ngram2", dict(prefix_allowed_tokens_fn=small_fn, no_repeat_ngram_size=2))])
def test_greedy_with_logits_processors_matches_reference_generate(setup, ctrl, name, kw):
    """HF greedy applies the repetition penalty to the RAW logits, then the n-gram ban / prefix constraint, pads finished rows
    and stops when all rows have finished — ids must equal the reference's for every combination (the 6-token `small_fn`
    vocabulary forces repeats so that the processors change the outcome)."""
    from emu_b200 import generation
    gold, sd, emb = setup
    toks = generation.greedy_search(OracleEngine(sd), emb, gold["gen_attention_mask"], 12, 2, 32000, min_length=1,
                                    check_every=4, **kw)
    assert torch.equal(toks, ctrl[name]), (toks.tolist(), ctrl[name].tolist())


def test_processors_change_the_outcome(ctrl):
    """guards the fixture itself: the repeat-forcing cases must differ from their unprocessed twins"""
    assert not torch.equal(ctrl["greedy_small"], ctrl["greedy_small_rep"])
    assert ctrl["greedy_small"].shape != ctrl["greedy_small_ngram2"].shape
    assert ctrl["beam3_small"].shape != ctrl["beam3_small_rep_ngram"].shape


@pytest.mark.parametrize("name,nb,kw", [
    ("beam3_rep_ngram", 3, dict(length_penalty=1.0, repetition_penalty=1.4, no_repeat_ngram_size=2)),
    ("beam3_prefix", 3, dict(length_penalty=0.0, prefix_allowed_tokens_fn=prefix_fn)),
    ("beam3_small", 3, dict(length_penalty=1.0, prefix_allowed_tokens_fn=small_fn)),
    ("beam3_small_rep_ngram", 3, dict(length_penalty=1.0, prefix_allowed_tokens_fn=small_fn, repetition_penalty=1.4,
                                      no_repeat_ngram_size=2)),
    ("beam4_ret3", 4, dict(length_penalty=-1, num_return_sequences=3))])
def test_beam_search_control_matches_reference_generate(setup, ctrl, name, nb, kw):
    """beam search with the processors on the log-probabilities (HF `_beam_search` order), a prefix constraint, and
    num_return_sequences = 3 (the n best hypotheses of every prompt, best first, [B * n, T])"""
    from emu_b200 import generation
    gold, sd, emb = setup
    toks = generation.beam_search(OracleEngine(sd, max_batch=8), emb, gold["gen_attention_mask"], nb, 12, 2, 32000,
                                  min_length=1, **kw)
    assert torch.equal(toks, ctrl[name]), (toks.tolist(), ctrl[name].tolist())


@pytest.mark.parametrize("name,B,nb,seed,kw", [
    ("beamsample3", 2, 3, 77, dict(temperature=0.8, top_p=0.9, length_penalty=1.0)),
    ("beamsample2_topk", 1, 2, 5, dict(top_k=8, length_penalty=-1)),
    # the chat demo's default knobs with "do sample" ticked: fewer live candidates (top_k = 3) than the 2 * num_beams drawn
    ("beamsample5_demo", 1, 5, 9, dict(top_k=3, top_p=0.9, temperature=0.7, length_penalty=1.0))])
def test_beam_sample_matches_reference_generate(setup, ctrl, name, B, nb, seed, kw):
    """do_sample with num_beams > 1 (what the chat demo runs with its defaults when do_sample is ticked,
    Emu2/demo/backend/pytorch_model/backend.py:196-214): processors -> warpers -> + running score -> 2*beams candidates drawn
    without replacement.  On the CPU the draws come from the same torch generator stream as HF's, so with the same seed the
    ids are the reference's, token for token."""
    from emu_b200 import generation
    gold, sd, emb = setup
    torch.manual_seed(seed)
    toks = generation.beam_search(OracleEngine(sd), emb[:B], gold["gen_attention_mask"][:B], nb, 12, 2, 32000, min_length=1,
                                  do_sample=True, **kw)
    assert torch.equal(toks, ctrl[name]), (toks.tolist(), ctrl[name].tolist())


@pytest.mark.parametrize("check_every", [1, 3, 5, 8, 0])
def test_beam_search_result_does_not_depend_on_when_the_host_looks(setup, ctrl, check_every):
    """The search below finishes after 9 of 12 steps.  The host polls `done` every `check_every` steps; the steps launched
    after the search has finished are no-ops on the beam state and must not change which sequence plane is read back
    (regression: an odd number of such steps used to return the hypotheses of the step BEFORE the last one)."""
    from emu_b200 import generation
    gold, sd, emb = setup
    toks = generation.beam_search(OracleEngine(sd, max_batch=8), emb, gold["gen_attention_mask"], 3, 12, 2, 32000,
                                  min_length=1, length_penalty=1.0, prefix_allowed_tokens_fn=small_fn,
                                  repetition_penalty=1.4, no_repeat_ngram_size=2, check_every=check_every)
    assert torch.equal(toks, ctrl["beam3_small_rep_ngram"])


def test_sampling_with_processors_and_return_sequences(setup):
    """do_sample with num_beams = 1: the processors run before the warpers (HF `_sample`), and num_return_sequences = n
    yields n continuations per prompt from ONE prefill ([B * n, T], prompt-major like HF's expanded batch)."""
    from emu_b200 import generation
    gold, sd, emb = setup
    g = torch.Generator().manual_seed(11)
    toks = generation.sample_search(OracleEngine(sd, max_batch=8), emb, gold["gen_attention_mask"], 10, 2, 32000, min_length=10,
                                    temperature=0.9, top_k=4, generator=g, prefix_allowed_tokens_fn=small_fn,
                                    no_repeat_ngram_size=2, num_return_sequences=3)
    assert toks.shape == (6, 10)
    assert set(toks.flatten().tolist()) <= set(SMALL)                    # the prefix constraint held at every step
    for row in toks.tolist():                                            # ... and no bigram occurs twice
        grams = list(zip(row, row[1:]))
        assert len(grams) == len(set(grams))
    assert len({tuple(r) for r in toks[:3].tolist()}) > 1                # the n continuations of one prompt are different draws
    with pytest.raises(ValueError):
        generation.sample_search(OracleEngine(sd, max_batch=4), emb, gold["gen_attention_mask"], 4, 2, 32000,
                                 num_return_sequences=3)


def _literal_contrastive(sd, emb, mask, max_new, k, alpha, eos, pad, min_length):
    """transformers 4.31 `contrastive_search` + `_ranking_fast` restated WITHOUT a cache (every candidate is a full forward of
    prompt + generated + candidate), pads excluded from the similarity — the checker for generation.contrastive_search."""
    B = emb.shape[0]
    E = sd["decoder.lm.model.embed_tokens.weight"]

    def fwd(e, m):
        return O.llama_forward(sd, e, m, layers=L, heads=NH, position_ids=O.hf_position_ids(m))
    seq, m = emb.clone(), mask.clone()
    h = fwd(seq, m)
    ctx, logits = h.float(), O.lm_logits(sd, h[:, -1]).float()
    out, unfinished = [], torch.ones(B, dtype=torch.long)
    for step in range(max_new):
        if step < min_length:
            logits[:, eos] = float("-inf")
        top_p, top_ids = torch.topk(torch.softmax(logits, -1), k)
        m1 = torch.cat((m, torch.ones(B, 1, dtype=m.dtype)), 1)
        nh, nl = [], []
        for j in range(k):
            hh = fwd(torch.cat((seq, F.embedding(top_ids[:, j], E).unsqueeze(1)), 1), m1)
            nh.append(hh[:, -1].float())
            nl.append(O.lm_logits(sd, hh[:, -1]).float())
        nh, nl = torch.stack(nh, 1), torch.stack(nl, 1)                                   # [B, k, H], [B, k, V]
        cos = torch.matmul(nh / nh.norm(dim=2, keepdim=True), (ctx / ctx.norm(dim=2, keepdim=True)).transpose(1, 2))
        cos = cos.masked_fill(m[:, None, :] == 0, float("-inf"))
        sel = ((1.0 - alpha) * top_p - alpha * cos.max(-1)[0]).argmax(-1)
        r = torch.arange(B)
        nxt = top_ids[r, sel] * unfinished + pad * (1 - unfinished)
        out.append(nxt)
        seq = torch.cat((seq, F.embedding(top_ids[r, sel], E).unsqueeze(1)), 1)
        ctx, logits, m = torch.cat((ctx, nh[r, sel][:, None]), 1), nl[r, sel], m1
        unfinished = unfinished * (nxt != eos).long()
        if int(unfinished.max()) == 0:
            break
    return torch.stack(out, 1)


@pytest.mark.parametrize("k,alpha", [(4, 0.6), (2, 0.3)])
def test_contrastive_search_matches_literal_restatement(setup, k, alpha):
    """penalty_alpha + top_k with one beam: the engine form (candidates as cache rows, expand -> one decode step -> collapse)
    against the cache-less literal algorithm.  PARITY UNPINNED for this strategy: transformers 5.5 (installed) no longer ships
    contrastive search, so the reference itself cannot produce a golden here."""
    from emu_b200 import generation
    gold, sd, emb = setup
    mask = gold["gen_attention_mask"]
    want = _literal_contrastive(sd, emb, mask, 8, k, alpha, 2, 32000, 1)
    got = generation.contrastive_search(OracleEngine(sd, max_batch=8), emb, mask, 8, 2, 32000, top_k=k, penalty_alpha=alpha,
                                        min_length=1, check_every=4)
    assert torch.equal(got, want), (got.tolist(), want.tolist())
    plain = generation.greedy_search(OracleEngine(sd), emb, mask, 8, 2, 32000, min_length=1)
    assert not torch.equal(got, plain) or alpha < 0.5          # the degeneration penalty does steer the search
    with pytest.raises(ValueError):
        generation.contrastive_search(OracleEngine(sd, max_batch=4), emb, mask, 4, 2, 32000, top_k=4, penalty_alpha=0.6)


def test_generate_dispatch_follows_generation_mixin(setup, ctrl):
    """generation.generate picks the strategy from (do_sample, num_beams, penalty_alpha, top_k) like GenerationMixin 4.31 and
    forwards every knob; checked on cases whose ids are pinned to the reference above."""
    from emu_b200 import generation
    gold, sd, emb = setup
    mask = gold["gen_attention_mask"]
    common = dict(max_new_tokens=12, eos_token_id=2, pad_token_id=32000, min_length=1)
    eng = lambda: OracleEngine(sd, max_batch=8)
    assert torch.equal(generation.generate(eng(), emb, mask, num_beams=1, **common), gold["gen_ids_greedy"])
    assert torch.equal(generation.generate(eng(), emb[:1], mask[:1], num_beams=5, length_penalty=-1, **common),
                       gold["gen_ids_beam5"])
    assert torch.equal(generation.generate(eng(), emb, mask, num_beams=1, repetition_penalty=1.5,
                                           prefix_allowed_tokens_fn=small_fn, **common), ctrl["greedy_small_rep"])
    assert torch.equal(generation.generate(eng(), emb, mask, num_beams=4, length_penalty=-1, num_return_sequences=3,
                                           **common), ctrl["beam4_ret3"])
    torch.manual_seed(77)
    assert torch.equal(generation.generate(eng(), emb, mask, num_beams=3, do_sample=True, temperature=0.8, top_p=0.9,
                                           **common), ctrl["beamsample3"])
    # penalty_alpha alone (top_k unset) is NOT contrastive search in HF: greedy
    assert torch.equal(generation.generate(eng(), emb, mask, num_beams=1, penalty_alpha=0.6, **common), gold["gen_ids_greedy"])
    c = generation.generate(eng(), emb, mask, num_beams=1, penalty_alpha=0.6, top_k=4, **common)
    assert torch.equal(c, generation.contrastive_search(eng(), emb, mask, 12, 2, 32000, top_k=4, penalty_alpha=0.6, min_length=1))
    with pytest.raises(ValueError):
        generation.generate(eng(), emb, mask, num_beams=1, num_return_sequences=2, **common)
    with pytest.raises(ValueError):
        generation.generate(eng(), emb, mask, num_beams=2, num_return_sequences=3, **common)


# ---- BASELINE configs[0] in miniature: Emu1 image -> text (Emu1/inference.py:66-80 -> Emu.generate) ------------------------
class Emu1OracleEngine(OracleEngine):
    """OracleEngine + the Emu1 image path (pre-norm EVA ViT -> ln_visual, Causal-Former), same surface as _lib.Engine"""

    def __init__(self, sd, vis, max_batch=8):
        super().__init__(sd, max_batch)
        self.vis = vis
        self.cfg.llm_vocab = sd["decoder.lm.lm_head.weight"].shape[0]

    def vit_forward(self, image, n_query, pool=True):
        v, dt = self.vis, self.sd["ln_visual.weight"].dtype
        feats = O.vit_forward_features(self.sd, image.to(dt), patch=v["patch_size"], num_heads=v["width"] // v["head_width"],
                                       layers=v["layers"], postnorm=False)
        return F.layer_norm(feats, (v["width"],), self.sd["ln_visual.weight"], self.sd["ln_visual.bias"], 1e-6)

    def cformer_forward(self, feats, n_queries, out_dim):
        from helpers import emu1_t5_cfg
        from oracle import t5_oracle as T
        return T.causal_former(self.sd, feats, emu1_t5_cfg())

    def llm_embed(self, ids):
        return F.embedding(ids.long(), self.sd["decoder.lm.model.embed_tokens.weight"])


def _emu1_model(dtype):
    from types import SimpleNamespace as NS
    from helpers import EMU1_VIS, StubTokenizer, emu1_state_dict
    from emu_b200.emu1.modeling_emu import Emu
    sd = {k: v.to(dtype) for k, v in emu1_state_dict(EMU1_VIS).items()}
    m = Emu.__new__(Emu)                      # host code only: the engine behind it is the CPU test double
    m.engine, m.device_ = Emu1OracleEngine(sd, EMU1_VIS), torch.device("cpu")
    m.decoder, m.n_causal, m.hidden = NS(tokenizer=StubTokenizer()), 8, 256
    return m, sd


@pytest.mark.parametrize("name,kw", [("greedy", dict(num_beams=1)), ("beam3", dict(num_beams=3, length_penalty=0.0)),
                                     ("beam3_lp1_ret2", dict(num_beams=3, length_penalty=1.0, num_return_sequences=2))])
def test_emu1_generate_matches_reference(name, kw):
    """`Emu.generate` of the UNMODIFIED Emu1 reference, run the way it runs itself (bf16 model under bf16 autocast; two
    left-padded prompts of different length with one image each, tests/golden/gen_golden_emu1_generate.py), against this repo's
    Emu1 host code over the oracle in the same dtype policy.  Random-init logits are nearly flat and the reference's bf16
    kernels need not round like the oracle's, so: every row must reproduce the reference ids up to its first near-tie flip, and
    where it departs, the token it chose must be within the bf16 noise margin of the best one under the fp32 oracle."""
    gold = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "emu1_generate_tiny.pt"))
    m, sd = _emu1_model(torch.bfloat16)
    ids = m.generate_from_ids(gold["input_ids"], gold["attention_mask"], image=gold["image"], max_new_tokens=12, min_length=1,
                              eos_token_id=2, pad_token_id=32000, **kw)
    ref = gold["ids_" + name]
    assert ids.shape == ref.shape, (ids.shape, ref.shape)
    if torch.equal(ids, ref):
        return
    m32, sd32 = _emu1_model(torch.float32)
    emb = m32.engine.llm_embed(gold["input_ids"])
    f = m32.encode_image(gold["image"])
    emb[gold["input_ids"] == 32003] = f.reshape(-1, f.shape[-1])
    n = ids.shape[0] // gold["input_ids"].shape[0]
    emb, mask = emb.repeat_interleave(n, 0), gold["attention_mask"].repeat_interleave(n, 0)
    for b in range(ids.shape[0]):
        diff = (ids[b] != ref[b]).nonzero()
        if diff.numel() == 0:
            continue
        t = int(diff[0])
        assert t >= 1, "row %d departs from the reference at the first token" % b
        e = torch.cat((emb[b:b + 1], m32.engine.llm_embed(ref[b:b + 1, :t])), 1)
        mk = torch.cat((mask[b:b + 1], torch.ones(1, t, dtype=mask.dtype)), 1)
        h = O.llama_forward(sd32, e, mk, layers=L, heads=NH, position_ids=O.hf_position_ids(mk))
        lp = torch.log_softmax(O.lm_logits(sd32, h[:, -1]).float(), -1)[0]
        assert float(lp[ref[b, t]] - lp[ids[b, t]]) <= 3e-2 * float(lp.abs().max()), (name, b, t, ids[b], ref[b])


# ---- EmuModel.generate with an image AND a video in the prompt (Emu2/emu/emu.py:197-211) ------------------------------------
def _emu2_model():
    """EmuModel's host code (tokens -> embeddings -> <image> / [gIMG] splice -> strategy) over the CPU test double"""
    from types import SimpleNamespace as NS
    from helpers import StubTokenizer
    from emu_b200.emu2.emu import EmuModel
    sd = make_emu2_state_dict()
    m = EmuModel.__new__(EmuModel)
    eng = OracleEngine(sd, max_batch=8)
    eng.llm_embed = lambda ids: F.embedding(ids.long(), sd["decoder.lm.model.embed_tokens.weight"])
    m.engine, m.device_ = eng, torch.device("cpu")
    m.decoder, m.n_query, m.v_query = NS(tokenizer=StubTokenizer()), 4, 4
    m.encode_image = lambda image, n_query=None: O.encode_image(sd, image, patch=14, num_heads=4, layers=2, n_query=n_query or 4)
    m._project_up = lambda x: F.linear(x, sd["project_up.weight"])
    return m


@pytest.mark.parametrize("name,kw", [("greedy", dict(num_beams=1)), ("beam3", dict(num_beams=3, length_penalty=1.0))])
def test_generate_with_image_and_video_matches_reference(name, kw):
    """one picture (4 <image> slots) and three video frames (3 x 4 [gIMG] slots) in one prompt: ids of the unmodified
    reference's EmuModel.generate (tests/golden/gen_golden_video.py)"""
    gold = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "emu2_tiny_video.pt"))
    assert int((gold["input_ids"] == 32003).sum()) == 4 and int((gold["input_ids"] == 32004).sum()) == 12
    ids = _emu2_model().generate_from_ids(gold["input_ids"], gold["attention_mask"], image=gold["image"], video=gold["video"],
                                          image_token_id=32003, video_token_id=32004, max_new_tokens=10, min_len=1,
                                          eos_token_id=2, pad_token_id=32000, **kw)
    assert torch.equal(ids, gold["ids_" + name]), (ids.tolist(), gold["ids_" + name].tolist())
