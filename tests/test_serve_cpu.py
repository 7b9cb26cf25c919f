"""CPU: the serving shell (emu_b200/serve.py) — the demo's HTTP contract (Emu2/demo/backend/pytorch_model/backend.py:110-226),
the form encoding its front end posts (Emu2/demo/frontend/libs/chat_frontend.py:96-127), and the request batching that
replaces the reference's semaphore-guarded pipeline pool.  The pipelines are stand-ins with the reference's call signature:
the shell is host code, the engine behind it has its own tests."""
import base64
import io
import json
import threading
import time

import pytest
import requests
from PIL import Image

from emu_b200 import serve


class FakeChat:
    """EmuChatGeneration's surface as the shell uses it: __call__(inputs=, **knobs) -> str, forward_batch, batch capacity"""

    def __init__(self, rows=20, delay=0.0):
        self.rows, self.delay = rows, delay
        self.calls = []                     # (n_requests, knobs) per generate call
        self.thread_ids = set()
        self.gate = None                    # an Event: the FIRST call waits on it (lets a test pile requests up behind it)

    def _describe(self, inputs):
        return " ".join(i if isinstance(i, str) else "<img %dx%d>" % i.size for i in inputs)

    def __call__(self, inputs, **knobs):
        return self.forward_batch([inputs], **knobs)[0]

    def forward_batch(self, batch_inputs, **knobs):
        self.thread_ids.add(threading.get_ident())
        self.calls.append((len(batch_inputs), dict(knobs)))
        if self.gate is not None and len(self.calls) == 1:
            assert self.gate.wait(30)
        if self.delay:
            time.sleep(self.delay)
        if any("boom" in self._describe(i) for i in batch_inputs):
            raise RuntimeError("decoder exploded")
        return ["echo[%d beams]: %s" % (knobs["num_beams"], self._describe(i)) for i in batch_inputs]

    def max_requests_per_batch(self, num_beams=5):
        return max(1, self.rows // num_beams)


class FakeGen:
    def __init__(self):
        self.calls = []

    def __call__(self, inputs, guidance_scale, num_inference_steps):
        return self.forward_batch([inputs], guidance_scale=guidance_scale, num_inference_steps=num_inference_steps)[0]

    def forward_batch(self, batch_inputs, guidance_scale, num_inference_steps):
        self.calls.append((len(batch_inputs), guidance_scale, num_inference_steps))
        time.sleep(0.05)
        outs = []
        for inputs in batch_inputs:
            shade = sum(len(i) for i in inputs if isinstance(i, str)) % 256
            outs.append(type("Out", (), {"image": Image.new("RGB", (16, 16), (shade, num_inference_steps, 7)),
                                         "nsfw_content_detected": None})())
        return outs


def wait_until(cond, timeout=30.0):
    t0 = time.time()
    while not cond():
        assert time.time() - t0 < timeout, "timed out"
        time.sleep(0.002)


def png_bytes(size=(8, 6), color=(10, 200, 30)):
    buf = io.BytesIO()
    Image.new("RGB", size, color).save(buf, format="PNG")
    return buf.getvalue()


@pytest.fixture()
def server():
    chat, gen = FakeChat(), FakeGen()
    srv = serve.EmuServer(chat=chat, generate=gen, host="127.0.0.1", port=0, max_wait_ms=0).start()
    yield srv, chat, gen, "http://127.0.0.1:%d" % srv.port
    srv.shutdown()


def post_chat(url, prompt, files=None, **data):
    """exactly what the demo front end sends (chat_frontend.py:110-127)"""
    return requests.post(url + "/v1/mmc", files=files, data=dict(log_id="t", prompt=json.dumps(prompt), **data), timeout=30)


def test_chat_contract_multipart_and_defaults(server):
    srv, chat, gen, url = server
    key = "[<IMAGE0>]"
    r = post_chat(url, [["IMAGE", key], ["TEXT", "what is this?"]], files={key: (key, io.BytesIO(png_bytes()), "image/png")},
                  do_sample=False, max_new_tokens=32, temperature=0.7, top_k=3, top_p=0.9, length_penalty=1, num_beams=5,
                  repetition_penalty=1.0)
    assert r.status_code == 200
    assert r.json() == {"code": 0, "data": "echo[5 beams]: <img 8x6> what is this?"}
    n, knobs = chat.calls[-1]
    assert n == 1 and knobs == dict(do_sample=False, max_new_tokens=32, temperature=0.7, top_k=3, top_p=0.9,
                                    length_penalty=1.0, num_beams=5, repetition_penalty=1.0)
    # text only -> requests sends application/x-www-form-urlencoded; absent fields take the reference's defaults
    r = post_chat(url, [["TEXT", "hello"]], do_sample="True")
    assert r.json()["code"] == 0
    n, knobs = chat.calls[-1]
    assert knobs == dict(serve.CHAT_DEFAULTS, do_sample=True)


def test_generation_contract_webp_payload(server):
    srv, chat, gen, url = server
    r = requests.post(url + "/v1/mmg", data=dict(log_id="g", prompt=json.dumps([["TEXT", "a cat"]]),
                                                 classifier_free_guidance=3.0, steps=25), timeout=30)
    body = r.json()
    assert body["code"] == 0
    im = Image.open(io.BytesIO(base64.b64decode(body["data"])))
    assert im.format == "WEBP" and im.size == (16, 16)
    assert gen.calls[-1] == (1, 3.0, 25)
    # a required field missing -> code -1 with the message, HTTP 200, server alive (backend.py:141-150)
    r = requests.post(url + "/v1/mmg", data=dict(log_id="g", prompt=json.dumps([["TEXT", "a cat"]])), timeout=30)
    assert r.status_code == 200 and r.json()["code"] == -1 and "classifier_free_guidance" in r.json()["data"]


def test_errors_answer_code_minus_one_and_the_server_survives(server):
    srv, chat, gen, url = server
    assert post_chat(url, [["TEXT", "boom"]]).json() == {"code": -1, "data": "decoder exploded"}
    assert post_chat(url, [["IMAGE", "nope"]]).json()["code"] == -1          # picture named in the prompt was not uploaded
    assert post_chat(url, [["TEXT", "still alive"]]).json()["code"] == 0
    assert requests.post(url + "/v1/nothing", data={}, timeout=30).status_code == 404


def test_concurrent_requests_share_generate_calls():
    """While the pipeline is busy, requests pile up; the next admission takes everything with the same knobs up to
    rows // num_beams, leaves the others in line, and every caller gets ITS answer.  The reference would have run 9 calls."""
    chat = FakeChat(rows=20)
    chat.gate = threading.Event()
    srv = serve.EmuServer(chat=chat, host="127.0.0.1", port=0, max_wait_ms=0).start()
    url = "http://127.0.0.1:%d" % srv.port
    worker = srv.chat_workers[0]
    answers = {}

    def client(i, beams):
        answers[i] = post_chat(url, [["TEXT", "request %d" % i]], num_beams=beams).json()
    first = threading.Thread(target=client, args=(0, 5))
    first.start()
    wait_until(lambda: len(chat.calls) == 1)                 # request 0 is inside the pipeline, alone
    rest = [threading.Thread(target=client, args=(i, 5 if i < 7 else 2)) for i in range(1, 9)]
    for i, t in enumerate(rest):
        t.start()
        wait_until(lambda: worker.q.qsize() == i + 1)        # queued in order 1..8
    chat.gate.set()
    for t in [first] + rest:
        t.join()
    srv.shutdown()
    for i in range(9):
        assert answers[i] == {"code": 0, "data": "echo[%d beams]: request %d" % (5 if i < 7 else 2, i)}
    assert [(n, k["num_beams"]) for n, k in chat.calls] == [(1, 5), (4, 5), (2, 5), (2, 2)]   # cap 20 // 5 = 4; other knobs last
    assert worker.batches == [1, 4, 2, 2]
    assert len(chat.thread_ids) == 1                                     # one thread owns the pipeline


def test_one_bad_request_does_not_fail_its_batch_mates():
    """a shared generate call that raises is retried request by request: only the culprit answers code -1"""
    chat = FakeChat(rows=20)
    chat.gate = threading.Event()
    srv = serve.EmuServer(chat=chat, host="127.0.0.1", port=0, max_wait_ms=0).start()
    url = "http://127.0.0.1:%d" % srv.port
    worker = srv.chat_workers[0]
    answers = {}
    texts = ["warm up", "fine one", "boom", "fine two"]

    def client(i):
        answers[i] = post_chat(url, [["TEXT", texts[i]]]).json()
    ts = [threading.Thread(target=client, args=(i,)) for i in range(4)]
    ts[0].start()
    wait_until(lambda: len(chat.calls) == 1)
    for i, t in enumerate(ts[1:]):
        t.start()
        wait_until(lambda: worker.q.qsize() == i + 1)
    chat.gate.set()
    for t in ts:
        t.join()
    srv.shutdown()
    assert answers[2] == {"code": -1, "data": "decoder exploded"}
    assert [answers[i]["code"] for i in (0, 1, 3)] == [0, 0, 0] and answers[3]["data"].endswith("fine two")
    assert [n for n, _ in chat.calls] == [1, 3, 1, 1, 1]                # alone | shared (raises) | one by one


def test_oversized_body_is_refused():
    srv = serve.EmuServer(chat=FakeChat(), host="127.0.0.1", port=0, max_body_bytes=1000).start()
    r = requests.post("http://127.0.0.1:%d/v1/mmc" % srv.port, data={"prompt": "x" * 5000}, timeout=30)
    srv.shutdown()
    assert r.status_code == 413


def test_max_wait_gathers_a_batch_from_an_idle_queue():
    chat = FakeChat(rows=8)
    # cap 8 // 2 = 4 > 3 requests: the batch closes when the window does; a long window keeps slow machines from splitting it
    w = serve.BatchingWorker(chat, "chat", lambda knobs: chat.max_requests_per_batch(knobs["num_beams"]), max_wait_ms=1500)
    w.start()
    out = {}
    ts = [threading.Thread(target=lambda i=i: out.__setitem__(i, w.submit(["q%d" % i], dict(num_beams=2)))) for i in range(3)]
    for t in ts:
        t.start()
        time.sleep(0.01)
    for t in ts:
        t.join()
    w.stop()
    assert out == {i: "echo[2 beams]: q%d" % i for i in range(3)} and w.batches == [3]


def test_parse_form_multipart_roundtrip():
    req = requests.Request("POST", "http://x/v1/mmc", files={"pic": ("pic", io.BytesIO(b"\x89PNG\r\n\x00\xff"), "image/png")},
                           data={"prompt": json.dumps([["TEXT", "héllo"]]), "top_k": 3}).prepare()
    fields, files = serve.parse_form(req.headers["Content-Type"], req.body)
    assert fields == {"prompt": json.dumps([["TEXT", "héllo"]]), "top_k": "3"} and files == {"pic": b"\x89PNG\r\n\x00\xff"}


# ---- tensor-parallel serving: rank 0 serves HTTP and announces every admitted batch, the other ranks follow (gloo, world 2) ----
def _tp_worker(rank, world, port, q):
    import os
    import sys
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    chat = FakeChat(rows=10)
    if rank > 0:
        serve.follow({"chat": chat})                       # returns when rank 0 releases it
        q.put((rank, [(n, k["num_beams"]) for n, k in chat.calls]))
    else:
        srv = serve.EmuServer(chat=chat, host="127.0.0.1", port=0, max_wait_ms=0, announce=serve.make_announce()).start()
        url = "http://127.0.0.1:%d" % srv.port
        key = "p"
        a = post_chat(url, [["IMAGE", key], ["TEXT", "one"]], files={key: (key, io.BytesIO(png_bytes()), "image/png")}).json()
        b = post_chat(url, [["TEXT", "two"]], num_beams=2).json()
        srv.shutdown()
        serve.release_followers()
        q.put((rank, [(n, k["num_beams"]) for n, k in chat.calls], a["data"], b["data"]))
    dist.destroy_process_group()


def test_tensor_parallel_followers_run_the_same_batches_gloo_world2():
    import os
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29850 + os.getpid() % 100
    ps = [ctx.Process(target=_tp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
    assert res[0][1] == res[1][1] == [(1, 5), (1, 2)]          # same calls, same order, on both ranks (pictures included)
    assert res[0][2] == "echo[5 beams]: <img 8x6> one" and res[0][3] == "echo[2 beams]: two"
