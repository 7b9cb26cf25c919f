"""Autoregressive generation control on top of the engine's prefill/decode steps.

Restates the two HF `GenerationMixin` strategies the reference drives through `lm.generate(inputs_embeds=...)`
(Emu2/emu/emu.py:213-229; Emu1/models/modeling_emu.py:162-179): greedy search and beam search
(num_beams=5, length_penalty=-1 are the reference defaults), plus sampling.  Token selection runs on the GPU; the
greedy loop never synchronises with the host inside the loop (token ids are chained device-to-device through the
CUDA-graphed decode step), and neither does beam search: its per-step hypothesis bookkeeping is a device kernel
(emu_beam_step) and the host reads the "finished" flag every few steps only.

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
    # one fresh seed per call, DRAWN from the generator so that its state advances like torch.multinomial's would:
    # two do_sample calls with the same prompt differ, a re-seeded generator reproduces
    seed = int(torch.randint(0, 2 ** 62, (1,), generator=generator, device="cpu" if generator is None else generator.device))
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


@torch.no_grad()
def beam_search(engine, inputs_embeds, attention_mask, num_beams, max_new_tokens, eos_token_id, pad_token_id,
                min_length=0, length_penalty=1.0, early_stopping=False, repetition_penalty=1.0, no_repeat_ngram_size=0,
                prefix_allowed_tokens_fn=None, penalty_on_logits=False, check_every=8):
    """HF (transformers >= 4.50 vectorised) beam search, decoder_prompt_len = 0 because generation is driven by
    inputs_embeds.  Returns the best finished hypothesis per batch row, new tokens only, padded.

    A step never leaves the device: engine.beam_topk (emu_beam_topk: log_softmax, logits processors, running score, top
    2*beams over beams x vocab) -> engine.beam_step (emu_beam_step: running / finished hypothesis bookkeeping, next tokens and
    KV-cache reorder indices) -> the CUDA-graphed decode step, chained through device buffers.  The host reads the `done`
    flag every `check_every` steps only (steps launched after the search has finished are no-ops on the beam state).
    `prefix_allowed_tokens_fn(batch_id, ids)` is a host callback by construction (Emu1/mm_eval/models/emu.py:97-109): when
    it is given, the running sequences are read back every step to build the allowed-token mask.

    With num_beams=1, early_stopping=True and penalty_on_logits=True this is HF greedy search with logits processors."""
    Bt = inputs_embeds.shape[0]
    dev = inputs_embeds.device
    nb = num_beams
    if Bt * nb > engine.cfg.llm_max_batch:
        raise ValueError("batch x num_beams = %d exceeds the engine's llm_max_batch = %d" % (Bt * nb, engine.cfg.llm_max_batch))
    V = engine.cfg.llm_vocab
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
    keep = 2 * nb
    need_prev = repetition_penalty != 1.0 or bool(no_repeat_ngram_size)
    cur_len = 0
    while True:
        live = st.running_seq[st.live(cur_len)]                       # [Bt, nb, max_length] tokens generated so far
        allowed = None
        if prefix_allowed_tokens_fn is not None:
            seqs = live[:, :, :cur_len].to("cpu", torch.int64)
            allowed = torch.zeros(Bt * nb, V, dtype=torch.uint8)
            for b in range(Bt):
                for k in range(nb):
                    ok = prefix_allowed_tokens_fn(b, seqs[b, k])
                    allowed[b * nb + k, torch.as_tensor(list(ok), dtype=torch.long)] = 1
            allowed = allowed.to(dev)
        topk_lp, topk_i = engine.beam_topk(logits, st.running_scores, Bt, nb, keep,
                                           ban_id=eos_token_id if cur_len < min_length else -1,
                                           prev_tokens=live.view(Bt * nb, max_length) if need_prev and cur_len > 0 else None,
                                           prev_len=cur_len, repetition_penalty=repetition_penalty,
                                           penalty_on_logits=penalty_on_logits, no_repeat_ngram=no_repeat_ngram_size or 0,
                                           allowed=allowed)
        engine.beam_step(st, topk_lp, topk_i, cur_len, eos_token_id, length_penalty, early_stopping)
        cur_len += 1
        if cur_len >= max_length or (check_every and cur_len % check_every == 0):
            if st.is_done():
                break
        engine.llm_decode(token_ids=st.next_tokens, beam_src=st.beam_src, logits=logits_buf, B=Bt * nb)
        logits = logits_buf
    return st.result(cur_len).to(dev)
