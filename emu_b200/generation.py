"""Autoregressive generation control on top of the engine's prefill/decode steps.

Restates the two HF `GenerationMixin` strategies the reference drives through `lm.generate(inputs_embeds=...)`
(Emu2/emu/emu.py:213-229; Emu1/models/modeling_emu.py:162-179): greedy search and beam search
(num_beams=5, length_penalty=-1 are the reference defaults), plus sampling.  Token selection runs on the GPU; the
greedy loop never synchronises with the host inside the loop (token ids are chained device-to-device through the
CUDA-graphed decode step), and neither does beam search: its per-step hypothesis bookkeeping is a device kernel
(emu_beam_step) and the host reads the "finished" flag every few steps only.

When driven by inputs_embeds HF returns only the NEW tokens; so do these functions.

Strategies (the combinations `lm.generate` picks from do_sample / num_beams / penalty_alpha, and which callers reach them):
  greedy_search        num_beams=1                         README image captioning; + logits processors (repetition_penalty,
                                                           no_repeat_ngram_size, prefix_allowed_tokens_fn) like HF `_sample`
  sample_search        do_sample, num_beams=1              chat demo (Emu2/demo/backend/pytorch_model/backend.py:196-214)
  beam_search          num_beams>1 (reference default 5)   every default call; do_sample=True -> HF beam-sample (what the demo
                                                           runs when "do_sample" is ticked with its default num_beams=5)
  contrastive_search   penalty_alpha>0, top_k>1, 1 beam    Emu2/emu/emu.py:166,222 / Emu1 modeling_emu.py:110 forward the knob
`num_return_sequences` (Emu1 `num_captions`, modeling_emu.py:113,177) is honoured by beam search (n best hypotheses per prompt) and
sampling (n independent draws per prompt, one prefill).
"""
import torch


def _allowed_mask(prefix_allowed_tokens_fn, seqs, V, device):
    """PrefixConstrainedLogitsProcessor's mask: seqs [rows_outer, rows_inner, cur_len] (host int64) -> uint8 [rows, V], 1 = allowed.
    The callback is host Python by construction (Emu1/mm_eval/models/emu.py:97-109), so this path reads the tokens back."""
    Bt, nb = seqs.shape[0], seqs.shape[1]
    allowed = torch.zeros(Bt * nb, V, dtype=torch.uint8)
    for b in range(Bt):
        for k in range(nb):
            ok = prefix_allowed_tokens_fn(b, seqs[b, k])
            allowed[b * nb + k, torch.as_tensor(list(ok), dtype=torch.long)] = 1
    return allowed.to(device)


def _finish_rows(toks, eos_token_id, pad_token_id):
    """HF `_sample` output form: after a row's first EOS every later token is pad; trimmed to the step at which the last
    row finished (the loop may have run a few steps past it between two host checks)."""
    is_eos = toks == eos_token_id
    after = (is_eos.cumsum(1) - is_eos.long()) > 0
    toks = toks.masked_fill(after, pad_token_id)
    if bool(is_eos.any(1).all()):
        last = int((is_eos.float().argmax(1)).max()) + 1
        toks = toks[:, :last]
    return toks


@torch.no_grad()
def greedy_search(engine, inputs_embeds, attention_mask, max_new_tokens, eos_token_id, pad_token_id, min_length=0,
                  check_every=16, repetition_penalty=1.0, no_repeat_ngram_size=0, prefix_allowed_tokens_fn=None):
    """HF greedy: argmax; finished rows emit pad; stop when every row has produced EOS.  With a repetition penalty, an n-gram
    ban or a prefix constraint the per-step argmax goes through the processor kernels (`_greedy_processed`)."""
    if repetition_penalty != 1.0 or no_repeat_ngram_size or prefix_allowed_tokens_fn is not None:
        return _greedy_processed(engine, inputs_embeds, attention_mask, max_new_tokens, eos_token_id, pad_token_id, min_length,
                                 check_every, repetition_penalty, no_repeat_ngram_size or 0, prefix_allowed_tokens_fn)
    B = inputs_embeds.shape[0]
    dev = inputs_embeds.device
    engine.llm_reset()
    _, logits = engine.llm_prefill(inputs_embeds, attention_mask, hf_positions=True, want_logits=True)
    if min_length > 0:
        logits[:, eos_token_id] = float("-inf")
    out = torch.empty(max_new_tokens, B, dtype=torch.int32, device=dev)
    # two fixed token buffers used alternately, so the engine sees only two argument sets (CUDA graphs / plans are
    # keyed by their pointers); each step's ids are also copied into `out` device-to-device
    ping = [torch.empty(B, dtype=torch.int32, device=dev), torch.empty(B, dtype=torch.int32, device=dev)]
    ping[0].copy_(logits.argmax(-1))
    out[0].copy_(ping[0])
    n_done = 1
    for step in range(1, max_new_tokens):
        ban = eos_token_id if step < min_length else -1
        src, dst = ping[(step - 1) & 1], ping[step & 1]
        engine.llm_decode(token_ids=src, next_ids=dst, ban_id=ban, B=B)
        out[step].copy_(dst)
        n_done = step + 1
        if check_every and step % check_every == 0:
            if bool((out[:n_done] == eos_token_id).any(0).all()):
                break
    return _finish_rows(out[:n_done].t().to(torch.int64), eos_token_id, pad_token_id)


def _greedy_processed(engine, inputs_embeds, attention_mask, max_new_tokens, eos_token_id, pad_token_id, min_length,
                      check_every, repetition_penalty, no_repeat_ngram_size, prefix_allowed_tokens_fn):
    """Greedy search with HF's logits processors, in `_sample`'s order: RepetitionPenalty on the RAW logits -> NoRepeatNGram ->
    MinLength -> PrefixConstrained -> argmax.  The processors and the selection are the device kernels of emu_beam_topk run
    with one beam (log_softmax is monotone per row, so its top-1 is the processed logits' argmax); `input_ids` for the
    processors is what HF sees when generation is driven by inputs_embeds: the tokens generated so far, pads included."""
    B = inputs_embeds.shape[0]
    dev = inputs_embeds.device
    V = engine.cfg.llm_vocab
    engine.llm_reset()
    _, logits = engine.llm_prefill(inputs_embeds, attention_mask, hf_positions=True, want_logits=True)
    logits_buf = torch.empty_like(logits)
    prev = torch.full((B, max_new_tokens), pad_token_id, dtype=torch.int32, device=dev)
    zero_run = torch.zeros(B, 1, dtype=torch.float32, device=dev)
    finished = torch.zeros(B, dtype=torch.bool, device=dev)
    tok = torch.empty(B, dtype=torch.int32, device=dev)
    pad_t = torch.full((B,), pad_token_id, dtype=torch.int32, device=dev)
    n_done = 0
    for step in range(max_new_tokens):
        allowed = None
        if prefix_allowed_tokens_fn is not None:
            allowed = _allowed_mask(prefix_allowed_tokens_fn, prev[:, None, :step].to("cpu", torch.int64), V, dev)
        _, idx = engine.beam_topk(logits, zero_run, B, 1, 2, ban_id=eos_token_id if step < min_length else -1,
                                  prev_tokens=prev if step > 0 else None, prev_len=step,
                                  repetition_penalty=repetition_penalty, penalty_on_logits=True,
                                  no_repeat_ngram=no_repeat_ngram_size, allowed=allowed)
        nxt = torch.where(finished, pad_t, idx[:, 0].to(torch.int32))      # one beam: flat index == token id
        prev[:, step] = nxt
        finished |= nxt == eos_token_id
        n_done = step + 1
        if n_done == max_new_tokens:
            break
        if check_every and n_done % check_every == 0 and bool(finished.all()):
            break
        tok.copy_(nxt)
        engine.llm_decode(token_ids=tok, logits=logits_buf, B=B)
        logits = logits_buf
    return _finish_rows(prev[:, :n_done].to(torch.int64), eos_token_id, pad_token_id)


@torch.no_grad()
def sample_search(engine, inputs_embeds, attention_mask, max_new_tokens, eos_token_id, pad_token_id, min_length=0,
                  temperature=None, top_k=None, top_p=None, generator=None, repetition_penalty=1.0, no_repeat_ngram_size=0,
                  prefix_allowed_tokens_fn=None, num_return_sequences=1):
    """Multinomial sampling, HF `_sample` order: logits processors (repetition penalty on the raw logits, n-gram ban, min
    length, prefix constraint) -> warpers temperature -> top_k -> top_p -> one draw per row.  num_return_sequences = n draws
    n independent continuations per prompt (HF expands the inputs n times; here each prompt is prefilled once and its cache
    row mapped to n rows); returns [B * n, T]."""
    B0 = inputs_embeds.shape[0]
    n = max(1, int(num_return_sequences or 1))
    B = B0 * n
    dev = inputs_embeds.device
    V = engine.cfg.llm_vocab
    if B > engine.cfg.llm_max_batch:
        raise ValueError("batch x num_return_sequences = %d exceeds the engine's llm_max_batch = %d" % (B, engine.cfg.llm_max_batch))
    engine.llm_reset()
    _, logits = engine.llm_prefill(inputs_embeds, attention_mask, hf_positions=True, want_logits=True)
    if n > 1:
        engine.llm_expand(torch.arange(B, device=dev, dtype=torch.int32) // n, B)
        logits = logits.repeat_interleave(n, dim=0).contiguous()
    processed = repetition_penalty != 1.0 or bool(no_repeat_ngram_size) or prefix_allowed_tokens_fn is not None
    out = []
    finished = torch.zeros(B, dtype=torch.bool, device=dev)
    nxt32 = torch.empty(B, dtype=torch.int32, device=dev)
    logits_buf = torch.empty_like(logits)
    prev = torch.full((B, max_new_tokens), pad_token_id, dtype=torch.int32, device=dev) if processed else None
    zero_run = torch.zeros(B, 1, dtype=torch.float32, device=dev)
    # one fresh seed per call, DRAWN from the generator so that its state advances like torch.multinomial's would:
    # two do_sample calls with the same prompt differ, a re-seeded generator reproduces
    seed = int(torch.randint(0, 2 ** 62, (1,), generator=generator, device="cpu" if generator is None else generator.device))
    for step in range(max_new_tokens):
        ban = eos_token_id if step < min_length else -1
        if processed:
            # the processor kernels of emu_beam_topk (one beam per row) leave the processed log-probabilities in `logits`;
            # softmax(log_softmax(x) / T) == softmax(x / T), so the warpers + draw below see what HF's see
            allowed = None
            if prefix_allowed_tokens_fn is not None:
                allowed = _allowed_mask(prefix_allowed_tokens_fn, prev[:, None, :step].to("cpu", torch.int64), V, dev)
            engine.beam_topk(logits, zero_run, B, 1, 2, ban_id=ban, prev_tokens=prev if step > 0 else None, prev_len=step,
                             repetition_penalty=repetition_penalty, penalty_on_logits=True,
                             no_repeat_ngram=no_repeat_ngram_size or 0, allowed=allowed)
        # device-side step (emu_sample_tokens): warpers + multinomial draw in the library
        nxt = engine.sample_tokens(logits, temperature or 1.0, top_k or 0, 1.0 if top_p is None else top_p, ban, seed,
                                   step).long()
        nxt = torch.where(finished, torch.full_like(nxt, pad_token_id), nxt)
        out.append(nxt)
        if processed:
            prev[:, step] = nxt.to(torch.int32)
        finished |= nxt == eos_token_id
        if bool(finished.all()) or step == max_new_tokens - 1:
            break
        nxt32.copy_(nxt)
        engine.llm_decode(token_ids=nxt32, logits=logits_buf, B=B)
        logits = logits_buf
    return torch.stack(out, dim=1)


def _warp_log_probs(lp, temperature, top_k, top_p, min_keep):
    """HF's sampling warpers on [rows, V] scores, in GenerationMixin's order (TemperatureLogitsWarper -> TopKLogitsWarper ->
    TopPLogitsWarper); under beam search they keep at least `min_keep` = 2 tokens per row so a beam can always continue."""
    if temperature is not None and temperature != 1.0:
        lp = lp / temperature
    if top_k:
        k = min(max(int(top_k), min_keep), lp.shape[-1])
        kth = torch.topk(lp, k)[0][..., -1, None]
        lp = lp.masked_fill(lp < kth, float("-inf"))
    if top_p is not None and top_p < 1.0:
        s_sorted, s_idx = torch.sort(lp, descending=False)
        remove = s_sorted.softmax(-1).cumsum(-1) <= (1 - top_p)
        remove[..., -min_keep:] = False
        lp = lp.masked_fill(remove.scatter(1, s_idx, remove), float("-inf"))
    return lp


@torch.no_grad()
def beam_search(engine, inputs_embeds, attention_mask, num_beams, max_new_tokens, eos_token_id, pad_token_id,
                min_length=0, length_penalty=1.0, early_stopping=False, repetition_penalty=1.0, no_repeat_ngram_size=0,
                prefix_allowed_tokens_fn=None, penalty_on_logits=False, check_every=8, num_return_sequences=1,
                do_sample=False, temperature=None, top_k=None, top_p=None, generator=None):
    """HF (transformers >= 4.50 vectorised) beam search, decoder_prompt_len = 0 because generation is driven by
    inputs_embeds.  Returns the best `num_return_sequences` finished hypotheses per batch row ([B * n, T], best first), new
    tokens only, padded.

    A step never leaves the device: engine.beam_topk (emu_beam_topk: log_softmax, logits processors, running score, top
    2*beams over beams x vocab) -> engine.beam_step (emu_beam_step: running / finished hypothesis bookkeeping, next tokens and
    KV-cache reorder indices) -> the CUDA-graphed decode step, chained through device buffers.  The host reads the `done`
    flag every `check_every` steps only (steps launched after the search has finished are no-ops on the beam state).
    `prefix_allowed_tokens_fn(batch_id, ids)` is a host callback by construction (Emu1/mm_eval/models/emu.py:97-109): when
    it is given, the running sequences are read back every step to build the allowed-token mask.

    do_sample=True is HF's beam-sample (`_get_top_k_continuations`): the processor kernels leave the processed
    log-probabilities in place, the sampling warpers and "+ running beam score" follow, and the 2*beams candidates are DRAWN
    without replacement from softmax over beams x vocab instead of taken by top-k; the bookkeeping is the same device step.

    With num_beams=1, early_stopping=True and penalty_on_logits=True this is HF greedy search with logits processors."""
    Bt = inputs_embeds.shape[0]
    dev = inputs_embeds.device
    nb = num_beams
    n_ret = max(1, int(num_return_sequences or 1))
    if n_ret > nb:
        raise ValueError("num_return_sequences = %d has to be <= num_beams = %d" % (n_ret, nb))
    if Bt * nb > engine.cfg.llm_max_batch:
        raise ValueError("batch x num_beams = %d exceeds the engine's llm_max_batch = %d" % (Bt * nb, engine.cfg.llm_max_batch))
    V = engine.cfg.llm_vocab
    keep = 2 * nb
    # (do_sample with top_k < 2 * num_beams — the chat demo's defaults, top_k = 3 with 5 beams — leaves the first step fewer
    # live candidates than the 2 * num_beams it draws: torch.multinomial then returns zero-probability entries, whose
    # accumulated score is -inf; those beams are dead from the start, exactly as in HF)
    max_length = max_new_tokens
    # HF expands the inputs to batch x beams and prefills num_beams identical copies of every prompt; here each prompt is
    # prefilled ONCE and its cache row is then mapped to num_beams rows (same cache contents, 1 / num_beams of the work)
    engine.llm_reset()
    _, logits = engine.llm_prefill(inputs_embeds, attention_mask, hf_positions=True, want_logits=True)
    if nb > 1:
        engine.llm_expand(torch.arange(Bt * nb, device=dev, dtype=torch.int32) // nb, Bt * nb)
        logits = logits.repeat_interleave(nb, dim=0).contiguous()
    logits_buf = torch.empty_like(logits)
    st = engine.beam_state(Bt, nb, max_length, pad_token_id, dev)
    need_prev = repetition_penalty != 1.0 or bool(no_repeat_ngram_size)
    zero_run = torch.zeros(Bt, nb, dtype=torch.float32, device=dev) if do_sample else None
    cur_len = 0
    while True:
        live = st.running_seq[st.live(cur_len)]                       # [Bt, nb, max_length] tokens generated so far
        allowed = None
        if prefix_allowed_tokens_fn is not None:
            allowed = _allowed_mask(prefix_allowed_tokens_fn, live[:, :, :cur_len].to("cpu", torch.int64), V, dev)
        topk_lp, topk_i = engine.beam_topk(logits, zero_run if do_sample else st.running_scores, Bt, nb, keep,
                                           ban_id=eos_token_id if cur_len < min_length else -1,
                                           prev_tokens=live.view(Bt * nb, max_length) if need_prev and cur_len > 0 else None,
                                           prev_len=cur_len, repetition_penalty=repetition_penalty,
                                           penalty_on_logits=penalty_on_logits, no_repeat_ngram=no_repeat_ngram_size or 0,
                                           allowed=allowed)
        if do_sample:
            lp = _warp_log_probs(logits.view(Bt * nb, V), temperature, top_k, top_p, min_keep=2)
            acc = (lp.view(Bt, nb, V) + st.running_scores.view(Bt, nb, 1)).view(Bt, nb * V)
            pick = torch.multinomial(torch.softmax(acc, dim=-1), num_samples=keep, generator=generator)
            topk_lp, topk_i = torch.gather(acc, 1, pick).contiguous(), pick.to(torch.int32).contiguous()
        engine.beam_step(st, topk_lp, topk_i, cur_len, eos_token_id, length_penalty, early_stopping)
        cur_len += 1
        if cur_len >= max_length or (check_every and cur_len % check_every == 0):
            if st.is_done():
                break
        engine.llm_decode(token_ids=st.next_tokens, beam_src=st.beam_src, logits=logits_buf, B=Bt * nb)
        logits = logits_buf
    return st.result(cur_len, n_ret).to(dev)


@torch.no_grad()
def contrastive_search(engine, inputs_embeds, attention_mask, max_new_tokens, eos_token_id, pad_token_id, top_k,
                       penalty_alpha, min_length=0, repetition_penalty=1.0, no_repeat_ngram_size=0, check_every=16):
    """Contrastive search (Su et al. 2022) as transformers 4.31 — the version the reference pins, Emu2/requirements.txt:2 —
    runs it when `penalty_alpha > 0`, `top_k > 1`, one beam and no sampling (Emu2/emu/emu.py:166,222 forward the knob):

        candidates  = top_k tokens of softmax(processed logits)            (one look-ahead decode step for all of them)
        degeneration(v) = max_j cos(h_v, h_j) over the hidden states h_j of the context            (`_ranking_fast`)
        next        = argmax_v (1 - alpha) * p(v) - alpha * degeneration(v)

    `h` are the last hidden states (post final norm = HF hidden_states[-1]).  Engine mapping: the k candidates of every row
    are decoded as k cache rows (emu_llm_expand maps the row to k, one emu_llm_decode step returns their logits and hidden
    states), then the chosen candidate's row becomes the row again (emu_llm_expand with the selected indices) — the cache
    itself never moves, only the row table the decode attention reads through.  Left-pad positions of the prompt take no
    part in the similarity (they hold no token).  Parity: restated from the published algorithm; the installed transformers
    (5.5) no longer ships contrastive search, so this strategy is pinned only by the literal restatement in the CPU tests."""
    B = inputs_embeds.shape[0]
    k = int(top_k)
    dev = inputs_embeds.device
    if k < 2 or not penalty_alpha or penalty_alpha <= 0:
        raise ValueError("contrastive search needs top_k > 1 and penalty_alpha > 0")
    if B * k > engine.cfg.llm_max_batch:
        raise ValueError("batch x top_k = %d exceeds the engine's llm_max_batch = %d" % (B * k, engine.cfg.llm_max_batch))
    V, H = engine.cfg.llm_vocab, inputs_embeds.shape[-1]
    alpha = float(penalty_alpha)
    engine.llm_reset()
    hidden, logits = engine.llm_prefill(inputs_embeds, attention_mask, hf_positions=True, want_hidden=True, want_logits=True)
    N = hidden.shape[1]
    ctx = torch.zeros(B, N + max_new_tokens, H, dtype=torch.float32, device=dev)        # unit-norm context hidden states
    ctx[:, :N] = torch.nn.functional.normalize(hidden.float(), dim=-1)
    valid = torch.zeros(B, N + max_new_tokens, dtype=torch.bool, device=dev)
    valid[:, :N] = attention_mask.to(dev) != 0
    processed = repetition_penalty != 1.0 or bool(no_repeat_ngram_size)
    prev = torch.full((B, max_new_tokens), pad_token_id, dtype=torch.int32, device=dev)
    zero_run = torch.zeros(B, 1, dtype=torch.float32, device=dev)
    expand_idx = torch.arange(B * k, device=dev, dtype=torch.int32) // k
    row_base = torch.arange(B, device=dev) * k
    cand_logits = torch.empty(B * k, V, dtype=torch.float32, device=dev)
    cand_hidden = torch.empty(B * k, H, dtype=torch.bfloat16, device=dev)
    cand_tok = torch.empty(B * k, dtype=torch.int32, device=dev)
    finished = torch.zeros(B, dtype=torch.bool, device=dev)
    pad_t = torch.full((B,), pad_token_id, dtype=torch.int64, device=dev)
    n_done = 0
    for step in range(max_new_tokens):
        ban = eos_token_id if step < min_length else -1
        if processed or ban >= 0:
            engine.beam_topk(logits, zero_run, B, 1, 2, ban_id=ban, prev_tokens=prev if step > 0 else None, prev_len=step,
                             repetition_penalty=repetition_penalty, penalty_on_logits=True,
                             no_repeat_ngram=no_repeat_ngram_size or 0)
        top_p_, top_ids = torch.topk(torch.softmax(logits.float(), dim=-1), k)                # [B, k]
        # look-ahead: every candidate as its own cache row
        engine.llm_expand(expand_idx, B * k)
        cand_tok.copy_(top_ids.reshape(-1))
        engine.llm_decode(token_ids=cand_tok, logits=cand_logits, hidden=cand_hidden, B=B * k)
        h = torch.nn.functional.normalize(cand_hidden.float(), dim=-1).view(B, k, H)
        cos = torch.einsum("bkh,bsh->bks", h, ctx[:, :N + step])
        cos = cos.masked_fill(~valid[:, None, :N + step], float("-inf"))
        score = (1.0 - alpha) * top_p_ - alpha * cos.max(dim=-1)[0]
        sel = score.argmax(dim=-1)                                                           # [B]
        nxt = torch.where(finished, pad_t, top_ids.gather(1, sel[:, None])[:, 0])
        prev[:, step] = nxt.to(torch.int32)
        ctx[:, N + step] = h[torch.arange(B, device=dev), sel]
        valid[:, N + step] = True
        finished |= nxt == eos_token_id
        n_done = step + 1
        # the chosen candidate's row (its look-ahead step is the sequence's real step) becomes row b again
        engine.llm_expand((row_base + sel).to(torch.int32), B)
        logits = cand_logits.view(B, k, V)[torch.arange(B, device=dev), sel].contiguous()
        if n_done == max_new_tokens:
            break
        if check_every and n_done % check_every == 0 and bool(finished.all()):
            break
    return _finish_rows(prev[:, :n_done].to(torch.int64), eos_token_id, pad_token_id)


def generate(engine, inputs_embeds, attention_mask, max_new_tokens, eos_token_id, pad_token_id, do_sample=False, num_beams=1,
             min_length=0, length_penalty=1.0, repetition_penalty=1.0, penalty_alpha=None, top_k=None, top_p=None,
             temperature=None, no_repeat_ngram_size=0, prefix_allowed_tokens_fn=None, num_return_sequences=1,
             early_stopping=False, generator=None, check_every=None):
    """`lm.generate(inputs_embeds=...)` as the reference calls it (Emu2/emu/emu.py:213-229, Emu1/models/modeling_emu.py:162-179):
    picks the decoding strategy from the knobs the way GenerationMixin (transformers 4.31, the pinned version) does

        num_beams == 1, no sampling, penalty_alpha > 0 and top_k > 1   -> contrastive search
        num_beams == 1, no sampling                                    -> greedy search
        num_beams == 1, do_sample                                      -> multinomial sampling
        num_beams  > 1                                                 -> beam search / beam-sample (do_sample)

    and returns the new token ids [B * num_return_sequences, T]."""
    n_ret = max(1, int(num_return_sequences or 1))
    ngram = int(no_repeat_ngram_size or 0)
    rp = 1.0 if repetition_penalty is None else float(repetition_penalty)
    ce = {} if check_every is None else {"check_every": check_every}
    if num_beams is None or num_beams < 1:
        raise ValueError("num_beams has to be an integer >= 1")
    if num_beams == 1 and not do_sample:
        if n_ret != 1:
            raise ValueError("num_return_sequences has to be 1 when doing greedy / contrastive search, got %d" % n_ret)
        if penalty_alpha is not None and penalty_alpha > 0 and top_k is not None and top_k > 1:
            if prefix_allowed_tokens_fn is not None:
                raise NotImplementedError("prefix_allowed_tokens_fn under contrastive search")
            return contrastive_search(engine, inputs_embeds, attention_mask, max_new_tokens, eos_token_id, pad_token_id,
                                      top_k=top_k, penalty_alpha=penalty_alpha, min_length=min_length, repetition_penalty=rp,
                                      no_repeat_ngram_size=ngram, **ce)
        return greedy_search(engine, inputs_embeds, attention_mask, max_new_tokens, eos_token_id, pad_token_id,
                             min_length=min_length, repetition_penalty=rp, no_repeat_ngram_size=ngram,
                             prefix_allowed_tokens_fn=prefix_allowed_tokens_fn, **ce)
    if num_beams == 1:
        return sample_search(engine, inputs_embeds, attention_mask, max_new_tokens, eos_token_id, pad_token_id,
                             min_length=min_length, temperature=temperature, top_k=top_k, top_p=top_p, generator=generator,
                             repetition_penalty=rp, no_repeat_ngram_size=ngram,
                             prefix_allowed_tokens_fn=prefix_allowed_tokens_fn, num_return_sequences=n_ret)
    return beam_search(engine, inputs_embeds, attention_mask, num_beams, max_new_tokens, eos_token_id, pad_token_id,
                       min_length=min_length, length_penalty=length_penalty, early_stopping=early_stopping,
                       repetition_penalty=rp, no_repeat_ngram_size=ngram, prefix_allowed_tokens_fn=prefix_allowed_tokens_fn,
                       num_return_sequences=n_ret, do_sample=bool(do_sample), temperature=temperature, top_k=top_k,
                       top_p=top_p, generator=generator, **ce)
