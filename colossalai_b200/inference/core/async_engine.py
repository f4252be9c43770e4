"""Asyncio front-end of the engine: request streams + a background stepping loop.
Parity: reference `colossalai/inference/core/async_engine.py:29-332` (`RequstStream`, `Tracer`,
`AsyncInferenceEngine.{add_request,generate,abort,start_background_loop}`)."""
from __future__ import annotations

import asyncio
from functools import partial
from typing import AsyncIterator, Dict, Iterable, List, Optional, Set, Tuple

from ...logging import get_dist_logger
from .engine import InferenceEngine

__all__ = ["AsyncInferenceEngine", "RequstStream", "Tracer"]
logger = get_dist_logger(__name__)


def _on_loop_done(task: asyncio.Task, tracker: "Tracer") -> None:
    msg = "Engine background task failed; see the traceback above."
    try:
        try:
            task.result()
        except asyncio.CancelledError:
            return
        except Exception as exc:
            raise RuntimeError(msg) from exc
        raise RuntimeError("Engine loop finished unexpectedly.")
    except Exception as exc:
        tracker.propagate_exception(exc)
        raise exc


class RequstStream:
    """One in-flight request: an awaitable slot that receives the finished `Sequence`.  (The reference keeps the same
    misspelt public name.)"""

    def __init__(self, request_id: int) -> None:
        self.request_id = request_id
        self._future: asyncio.Future = asyncio.get_event_loop().create_future()

    def set_result(self, result) -> None:
        if not self._future.done():
            if isinstance(result, BaseException):
                self._future.set_exception(result)
            else:
                self._future.set_result(result)

    async def get_result(self):
        return await self._future

    @property
    def finished(self) -> bool:
        return self._future.done()


class Tracer:
    """Bookkeeping between the API coroutine side and the engine loop side."""

    def __init__(self) -> None:
        self._streams: Dict[int, RequstStream] = {}
        self._finished: asyncio.Queue = asyncio.Queue()
        self._new: asyncio.Queue = asyncio.Queue()
        self.new_requests_event: Optional[asyncio.Event] = None

    def __contains__(self, item) -> bool:
        return item in self._streams

    def init_event(self) -> None:
        self.new_requests_event = asyncio.Event()

    def propagate_exception(self, exc: Exception, request_id: Optional[int] = None) -> None:
        targets = [self._streams[request_id]] if request_id is not None else list(self._streams.values())
        for s in targets:
            s.set_result(exc)

    def process_finished_request(self, finished_request) -> None:
        rid = finished_request.request_id
        stream = self._streams.pop(rid, None)
        if stream is not None:
            stream.set_result(finished_request)

    def add_request(self, request_id: int, **engine_add_request_kwargs) -> RequstStream:
        if request_id in self._streams:
            raise KeyError(f"Request {request_id} already exists.")
        stream = RequstStream(request_id)
        self._new.put_nowait((stream, {"request_id": request_id, **engine_add_request_kwargs}))
        self.new_requests_event.set()
        return stream

    def abort_request(self, request_id: int, *, verbose: bool = False) -> None:
        if verbose:
            logger.info(f"Aborted request {request_id}.")
        self._finished.put_nowait(request_id)
        stream = self._streams.get(request_id)
        if stream is None or stream.finished:
            return
        stream.set_result(None)

    def get_new_requests(self) -> Tuple[List[dict], Set[int]]:
        new, finished = [], set()
        while not self._finished.empty():
            rid = self._finished.get_nowait()
            finished.add(rid)
            self._streams.pop(rid, None)
        while not self._new.empty():
            stream, req = self._new.get_nowait()
            if stream.request_id in finished:
                stream.set_result(None)
                continue
            self._streams[stream.request_id] = stream
            new.append(req)
        self.new_requests_event.clear()
        return new, finished

    async def wait_for_new_requests(self) -> None:
        await self.new_requests_event.wait()


class _AsyncInferenceEngine(InferenceEngine):
    async def async_step(self):
        """One scheduler+model step run in the default executor so the event loop stays responsive."""
        eng = self.engine
        loop = asyncio.get_event_loop()
        finished = await loop.run_in_executor(None, eng.step)
        # "still busy" must include the WAITING list: a step can retire every running sequence while requests admitted
        # at its start are still queued (decode has priority over prefill); looking at the batch buckets only would
        # park the loop on the new-request event with work pending - and with every client blocked on its own
        # unfinished request that event never fires
        rh = eng.request_handler
        return finished, rh.check_unfinished_reqs() or rh.total_requests_in_batch_bucket() > 0

    def add_single_request(self, request_id: int, prompt: str, prompt_token_ids=None, generation_config=None) -> None:
        self.engine.add_request(request_ids=request_id, prompts=prompt, prompts_token_ids=prompt_token_ids,
                                generation_config=generation_config)

    def abort_request(self, request_ids: Iterable[int]) -> None:
        for rid in request_ids:
            self.engine.request_handler.abort_sequence(rid)


class AsyncInferenceEngine:
    """Wraps an `InferenceEngine` behind `await engine.generate(request_id, prompt)`."""

    _engine_class = _AsyncInferenceEngine

    def __init__(self, start_engine_loop: bool = True, **kwargs) -> None:
        self.engine = self._engine_class(**kwargs)
        self.background_loop: Optional[asyncio.Future] = None
        self._background_loop_unshielded = None
        self.start_engine_loop = start_engine_loop
        self._request_tracer = Tracer()

    @property
    def background_loop_status(self) -> bool:
        return self.background_loop is not None and not self.background_loop.done()

    def start_background_loop(self) -> None:
        if self.background_loop_status:
            raise RuntimeError("Existing loop is running")
        self._request_tracer.init_event()
        self._background_loop_unshielded = asyncio.get_event_loop().create_task(self.run_engine_loop())
        self._background_loop_unshielded.add_done_callback(partial(_on_loop_done, tracker=self._request_tracer))
        self.background_loop = asyncio.shield(self._background_loop_unshielded)

    async def step(self) -> bool:
        new, finished = self._request_tracer.get_new_requests()
        for req in new:
            self.engine.add_single_request(req["request_id"], req.get("prompt"), req.get("prompt_token_ids"),
                                           req.get("generation_config"))
        if finished:
            self.engine.abort_request(finished)
        done, has_running = await self.engine.async_step()
        for seq in done:
            self._request_tracer.process_finished_request(seq)
        return has_running

    async def abort(self, request_id: int) -> None:
        if not self.background_loop_status:
            raise RuntimeError("Background loop is not running or launched correctly.")
        self._request_tracer.abort_request(request_id)

    async def run_engine_loop(self) -> None:
        running = False
        while True:
            if not running:
                await self._request_tracer.wait_for_new_requests()
            running = await self.step()
            await asyncio.sleep(0)

    async def add_request(self, request_id: int, prompt: Optional[str], prompt_token_ids: Optional[List[int]] = None,
                          generation_config=None) -> RequstStream:
        if not self.background_loop_status:
            if self.start_engine_loop:
                self.start_background_loop()
            else:
                raise RuntimeError("Background loop is not running.")
        return self._request_tracer.add_request(request_id, prompt=prompt, prompt_token_ids=prompt_token_ids,
                                                generation_config=generation_config)

    async def generate(self, request_id: int, prompt: Optional[str], prompt_token_ids: Optional[List[int]] = None,
                       generation_config=None) -> AsyncIterator[str]:
        try:
            stream = await self.add_request(request_id, prompt, prompt_token_ids, generation_config)
            seq = await stream.get_result()
            if seq is None:
                return
            yield self.engine.tokenizer_decode(seq) if hasattr(self.engine, "tokenizer_decode") else \
                self.engine.engine.tokenizer.decode(seq.output_token_id, skip_special_tokens=True)
        except (Exception, asyncio.CancelledError):
            self._request_tracer.abort_request(request_id)
            raise
