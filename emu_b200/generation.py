"""Autoregressive generation control on top of the engine's prefill/decode steps.

Restates the two HF `GenerationMixin` strategies the reference drives through `lm.generate(inputs_embeds=...)`
(Emu2/emu/emu.py:213-229; Emu1/models/modeling_emu.py:162-179): greedy search and beam search
(num_beams=5, length_penalty=-1 are the reference defaults), plus sampling.  Token selection runs on the GPU; the
greedy loop never synchronises with the host inside the loop (token ids are chained device-to-device through the
CUDA-graphed decode step), beam search synchronises once per step for the stopping test exactly like HF does.

When driven by inputs_embeds HF returns only the NEW tokens; so do these functions.
"""
import torch


@torch.no_grad()
def greedy_search(engine, inputs_embeds, attention_mask, max_new_tokens, eos_token_id, pad_token_id, min_length=0,
                  check_every=16):
    """HF greedy: argmax; finished rows emit pad; stop when every row has produced EOS."""
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
    toks = out[:n_done].t().to(torch.int64)  # [B, T]
    # HF semantics: after a row's first EOS every later token is pad; trim to the longest unfinished row + EOS
    is_eos = toks == eos_token_id
    after = (is_eos.cumsum(1) - is_eos.long()) > 0
    toks = toks.masked_fill(after, pad_token_id)
    if bool(is_eos.any(1).all()):
        last = int((is_eos.float().argmax(1)).max()) + 1
        toks = toks[:, :last]
    return toks


@torch.no_grad()
def sample_search(engine, inputs_embeds, attention_mask, max_new_tokens, eos_token_id, pad_token_id, min_length=0,
                  temperature=None, top_k=None, top_p=None, generator=None):
    """Multinomial sampling with HF's warper order: temperature -> top_k -> top_p."""
    B = inputs_embeds.shape[0]
    dev = inputs_embeds.device
    engine.llm_reset()
    _, logits = engine.llm_prefill(inputs_embeds, attention_mask, hf_positions=True, want_logits=True)
    out = []
    finished = torch.zeros(B, dtype=torch.bool, device=dev)
    nxt32 = torch.empty(B, dtype=torch.int32, device=dev)
    logits_buf = torch.empty_like(logits)
    seed = generator.initial_seed() if generator is not None else torch.initial_seed()
    for step in range(max_new_tokens):
        # device-side step (emu_sample_tokens): warpers + multinomial draw in the library
        nxt = engine.sample_tokens(logits, temperature or 1.0, top_k or 0, 1.0 if top_p is None else top_p,
                                   eos_token_id if step < min_length else -1, seed, step).long()
        nxt = torch.where(finished, torch.full_like(nxt, pad_token_id), nxt)
        out.append(nxt)
        finished |= nxt == eos_token_id
        if bool(finished.all()) or step == max_new_tokens - 1:
            break
        nxt32.copy_(nxt)
        engine.llm_decode(token_ids=nxt32, logits=logits_buf, B=B)
        logits = logits_buf
    return torch.stack(out, dim=1)


def _gather_beams(t, idx):
    while idx.dim() < t.dim():
        idx = idx.unsqueeze(-1)
    return torch.take_along_dim(t, idx, dim=1)


@torch.no_grad()
def beam_search(engine, inputs_embeds, attention_mask, num_beams, max_new_tokens, eos_token_id, pad_token_id,
                min_length=0, length_penalty=1.0, early_stopping=False, repetition_penalty=1.0):
    """HF (transformers >= 4.50 vectorised) beam search, decoder_prompt_len = 0 because generation is driven by
    inputs_embeds.  Returns the best finished hypothesis per batch row, new tokens only, padded."""
    Bt = inputs_embeds.shape[0]
    dev = inputs_embeds.device
    nb = num_beams
    if Bt * nb > engine.cfg.llm_max_batch:
        raise ValueError("batch x num_beams = %d exceeds the engine's llm_max_batch = %d" % (Bt * nb, engine.cfg.llm_max_batch))
    V = engine.cfg.llm_vocab
    max_length = max_new_tokens
    emb = inputs_embeds.repeat_interleave(nb, dim=0)
    mask = attention_mask.repeat_interleave(nb, dim=0) if attention_mask is not None else None
    engine.llm_reset()
    _, logits = engine.llm_prefill(emb, mask, hf_positions=True, want_logits=True)
    logits_buf = torch.empty_like(logits)

    keep = 2 * nb
    # The vocabulary-wide work of a step runs on the device (engine.beam_topk); what is left is bookkeeping on
    # [batch, 2*beams] / [batch, beams, len] tensors, a few hundred bytes.  It lives on the HOST: one 80-byte D2H of the
    # candidates (the step has to synchronise for the stopping test anyway) replaces ~40 tiny kernel launches per step.
    bdev = torch.device("cpu")
    top_mask = torch.cat((torch.ones(nb, dtype=torch.bool), torch.zeros(keep - nb, dtype=torch.bool))).to(bdev)
    running_seq = torch.full((Bt, nb, max_length), pad_token_id, dtype=torch.int64, device=bdev)
    sequences = running_seq.clone()
    running_scores = torch.zeros(Bt, nb, dtype=torch.float, device=bdev)
    running_scores[:, 1:] = -1e9
    beam_scores = torch.full((Bt, nb), -1e9, dtype=torch.float, device=bdev)
    is_finished = torch.zeros(Bt, nb, dtype=torch.bool, device=bdev)
    unsat = torch.ones(Bt, 1, dtype=torch.bool, device=bdev)
    run_beam_idx = torch.full((Bt, nb, max_length), -1, dtype=torch.int32, device=bdev)
    beam_idx_fin = run_beam_idx.clone()
    batch_off = (torch.arange(Bt, device=bdev) * nb).view(-1, 1)
    tok32 = torch.empty(Bt * nb, dtype=torch.int32, device=dev)
    src32 = torch.empty(Bt * nb, dtype=torch.int32, device=dev)

    cur_len = 0
    while True:
        # device-side step (emu_beam_topk): log_softmax + repetition penalty + EOS ban + running score + top-2*beams over
        # beams x vocab all happen in the library; only the [batch, 2*beams] bookkeeping below runs here
        prev = None
        if repetition_penalty != 1.0 and cur_len > 0:
            prev = running_seq[:, :, :cur_len].reshape(Bt * nb, cur_len).to(dev)
        topk_lp, topk_i = engine.beam_topk(logits, running_scores.to(dev), Bt, nb, keep,
                                           ban_id=eos_token_id if cur_len < min_length else -1, prev_tokens=prev,
                                           repetition_penalty=repetition_penalty)
        topk_lp, topk_i = topk_lp.to(bdev), topk_i.to(bdev)  # [batch, 2*beams]: the only per-step device->host traffic
        topk_beam = topk_i // V
        topk_ids = topk_i % V
        topk_run_bi = _gather_beams(run_beam_idx, topk_beam)
        topk_seq = _gather_beams(running_seq, topk_beam)
        topk_seq[:, :, cur_len] = topk_ids
        topk_run_bi[:, :, cur_len] = (topk_beam + batch_off).to(torch.int32)
        hits = (topk_ids == eos_token_id) | (cur_len + 1 >= max_length)

        # running beams for the next iteration
        run_lp = topk_lp + hits.float() * -1.0e9
        nxt_i = torch.topk(run_lp, k=nb)[1]
        running_seq = _gather_beams(topk_seq, nxt_i)
        running_scores = _gather_beams(run_lp, nxt_i)
        run_beam_idx = _gather_beams(topk_run_bi, nxt_i)

        # finished beams
        just_fin = hits & top_mask[None, :]
        fin_lp = topk_lp / ((cur_len + 1) ** length_penalty)
        full = torch.all(is_finished, dim=-1, keepdim=True) & (early_stopping is True)
        fin_lp = fin_lp + full.float() * -1.0e9
        fin_lp = fin_lp + (~unsat).float() * -1.0e9
        fin_lp = fin_lp + (~just_fin).float() * -1.0e9
        m_seq = torch.cat((sequences, topk_seq), dim=1)
        m_sc = torch.cat((beam_scores, fin_lp), dim=1)
        m_bi = torch.cat((beam_idx_fin, topk_run_bi), dim=1)
        m_fin = torch.cat((is_finished, just_fin), dim=1)
        sel = torch.topk(m_sc, k=nb)[1]
        sequences = _gather_beams(m_seq, sel)
        beam_scores = _gather_beams(m_sc, sel)
        beam_idx_fin = _gather_beams(m_bi, sel)
        is_finished = _gather_beams(m_fin, sel)

        src = run_beam_idx[:, :, cur_len].reshape(-1)
        cur_len += 1
        if early_stopping == "never" and length_penalty > 0.0:
            best_len = max_length
        else:
            best_len = cur_len
        best_running = running_scores[:, :1] / (best_len ** length_penalty)
        worst_fin = torch.where(is_finished, beam_scores.min(dim=1, keepdim=True)[0], -1.0e9)
        unsat = unsat & torch.any(best_running > worst_fin, dim=-1, keepdim=True)
        improvement = torch.any(unsat)
        open_beam = ~(torch.all(is_finished) & (early_stopping is True))
        valid = ~torch.all(hits)
        if not bool(improvement & open_beam & valid):
            break
        tok32.copy_(running_seq[:, :, cur_len - 1].reshape(-1))
        src32.copy_(src)
        engine.llm_decode(token_ids=tok32, beam_src=src32, logits=logits_buf, B=Bt * nb)
        logits = logits_buf

    best = sequences[:, 0, :]
    bi = beam_idx_fin[:, 0, :]
    gen_len = int(((bi + 1).bool()).sum(dim=1).max())
    return best[:, :gen_len].to(dev)
