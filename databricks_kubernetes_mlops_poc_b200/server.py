"""FastAPI serving layer: the reference's ``app/main.py`` with the new request-batching loop.

Same HTTP surface as the reference (``app/main.py:35-43``): ``POST /predict`` takes a JSON list of
``LoanApplicant`` rows and returns ``ModelOutput``; Swagger UI at ``/``; ``MODEL_DIRECTORY`` and
``SERVICE_NAME`` environment variables; the model is loaded once in ``lifespan`` (``:20-31``) and cleared
at shutdown; both log records (``type: InferenceData`` ``:60-69`` and ``type: ModelOutput`` ``:75-84``) keep
their schema because the reference's KQL dashboards query it.  Error behaviour is the reference's too:
type errors are FastAPI's 422, anything raised while scoring is a 500, an empty list is a 500.

What is new sits between the reference's lines 54 and 72:

* a request body of the regular shape goes from bytes to the 23 columns in one pass of the native parser
  (``ingest.py`` / ``csrc/json_rows.h``: float64 arrays + Arrow string buffers, no per-row Python objects); any other
  body is validated by pydantic-core with the same field rules and 422 behaviour as ``list[LoanApplicant]`` (no
  ``pd.DataFrame(list_of_models)`` either way);
* a micro-batcher collects concurrent requests for up to ``B200_BATCH_WINDOW_US`` microseconds (or
  ``B200_MAX_BATCH`` rows), dictionary-encodes them into one pinned staging slot, and scores the whole
  slot with ONE engine call (H2D + classifier kernel + outlier-forest kernel + D2H); batches are dealt round-robin
  to the GPUs listed in ``B200_DEVICES`` -- the reference instead blocks its event loop per request (``async def``
  calling blocking code, ``app/main.py:43,72``), so requests are strictly serialised there;
* the per-request drift scores (GPU, ``drift.py``) run concurrently with that on their own stream and thread;
* the two JSON log lines are produced on a logging thread, off the request's critical path.
"""

from __future__ import annotations

import asyncio
import json
import logging
import os
import queue
import threading
import time
import uuid
from concurrent.futures import ThreadPoolExecutor
from contextlib import asynccontextmanager
from typing import AsyncGenerator

import numpy as np
import pandas as pd
from fastapi import FastAPI, Request, Response

from .ingest import NativeRequestParser, parse_rows, rows_to_frame  # noqa: F401  (rows_to_frame re-exported)
from .schema import ALL_FEATURES, LoanApplicant, ModelOutput

ml_models: dict = {}


def _service_name() -> str:
    return os.environ.get("SERVICE_NAME", "credit-default-api")


parse_request = parse_rows  # the general validator (pydantic-core over the raw bytes); see ingest.py


_REQUEST_SCHEMA = {"required": True, "content": {"application/json": {"schema": {
    "title": "Data", "type": "array", "items": LoanApplicant.model_json_schema()}}}}


class _Pending:
    __slots__ = ("frame", "future", "loop", "n")

    def __init__(self, frame, future, loop):
        self.frame, self.future, self.loop, self.n = frame, future, loop, len(frame)


class MicroBatcher:
    """Cross-request batching in front of a replica's ``score`` (one worker thread per model).

    ``models`` is a list (one per GPU); consecutive batches go round-robin over it."""

    def __init__(self, models, max_rows: int = 65536, window_us: int = 200):
        self.models = list(models)
        self.max_rows = int(max_rows)
        self.window_s = window_us * 1e-6
        self.q: queue.Queue = queue.Queue()
        self.batches = 0
        self.rows = 0
        self._stop = False
        self._threads = [threading.Thread(target=self._run, args=(i,), daemon=True, name=f"b200-batcher-{i}")
                         for i in range(len(self.models))]
        for t in self._threads:
            t.start()

    async def score(self, frame: pd.DataFrame):
        """-> (proba1 (n,), is_outlier (n,) or None) for this request's rows."""
        loop = asyncio.get_running_loop()
        fut = loop.create_future()
        self.q.put(_Pending(frame, fut, loop))
        return await fut

    def close(self) -> None:
        self._stop = True
        for _ in self._threads:
            self.q.put(None)
        for t in self._threads:
            t.join(timeout=5)

    def _collect(self):
        first = self.q.get()
        if first is None:
            return None
        items, rows = [first], first.n
        deadline = time.perf_counter() + self.window_s
        while rows < self.max_rows:
            left = deadline - time.perf_counter()
            try:
                nxt = self.q.get(timeout=left) if left > 0 else self.q.get_nowait()
            except queue.Empty:
                break
            if nxt is None:
                self.q.put(None)
                break
            items.append(nxt)
            rows += nxt.n
        return items

    def _score_items(self, model, items):
        """One engine call for the whole batch -> per-item (proba, flags) parts."""
        frame = items[0].frame if len(items) == 1 else pd.concat([it.frame for it in items], ignore_index=True)
        # encode -> pinned slot -> H2D -> classifier kernel (+ outlier-forest kernel) -> D2H
        scorer = getattr(model, "score", None)
        proba, flags = scorer(frame) if scorer is not None else (model.predict_proba1(frame), None)
        self.batches += 1
        self.rows += len(frame)
        parts, off = [], 0
        for it in items:
            parts.append((proba[off:off + it.n], None if flags is None else flags[off:off + it.n]))
            off += it.n
        return parts

    def _run(self, idx: int) -> None:
        model = self.models[idx]
        while not self._stop:
            items = self._collect()
            if items is None:
                return
            try:
                parts = self._score_items(model, items)
                for it, part in zip(items, parts):
                    it.loop.call_soon_threadsafe(_resolve, it.future, part, None)
            except BaseException as e:  # surfaces as HTTP 500, like any model exception in the reference
                if len(items) == 1:
                    items[0].loop.call_soon_threadsafe(_resolve, items[0].future, None, e)
                    continue
                # the reference scores requests independently (app/main.py:72): a request the model rejects (a value
                # that overflows float32, NaN with the outlier forest attached ...) must fail ALONE -- re-score the
                # batch one request at a time and route each outcome to its own caller
                for it in items:
                    try:
                        part = self._score_items(model, [it])[0]
                        it.loop.call_soon_threadsafe(_resolve, it.future, part, None)
                    except BaseException as e1:
                        it.loop.call_soon_threadsafe(_resolve, it.future, None, e1)


def _resolve(fut, value, err):
    if fut.cancelled():
        return
    if err is not None:
        fut.set_exception(err)
    else:
        fut.set_result(value)


class _BoundedLogPool:
    """One logging thread with a BOUNDED backlog: each queued record holds its request's DataFrame, so an unbounded
    queue grows without limit when logging falls behind.  Above the backlog the record is written inline on the
    caller's thread (back-pressure, nothing is dropped: the log schema is an API for the reference's KQL queries)."""

    def __init__(self, backlog: int = 256):
        self._pool = ThreadPoolExecutor(max_workers=1, thread_name_prefix="b200-log")
        self._slots = threading.BoundedSemaphore(max(1, backlog))
        self.inline = 0

    def submit(self, fn, *args):
        if not self._slots.acquire(blocking=False):
            self.inline += 1
            fn(*args)
            return None

        def run():
            try:
                fn(*args)
            finally:
                self._slots.release()

        return self._pool.submit(run)


def _as_list(a: np.ndarray) -> list:
    """ndarray -> list; large float64 / int32 results go through the recycling list builder (``_pylists.py``)."""
    if len(a) >= 256 and a.ndim == 1 and a.dtype in (np.float64, np.int32):
        from ._pylists import ListBuilder

        b = ListBuilder(len(a))
        b.fill(0, a)
        return b.items
    return a.tolist()


def _log_record(kind: str, request_id: str, payload) -> None:
    logging.info(json.dumps({"service_name": _service_name(), "type": kind, "request_id": request_id, "data": payload}))


def create_app(model=None, loader=None) -> FastAPI:
    """Build the app.  ``model``: an already-built B200Model (tests); otherwise ``loader`` (default
    ``databricks_kubernetes_mlops_poc_b200.load_model``) is called in ``lifespan`` on ``MODEL_DIRECTORY``."""
    log_pool = _BoundedLogPool(int(os.environ.get("B200_LOG_BACKLOG", "256")))
    parser = NativeRequestParser()

    @asynccontextmanager
    async def lifespan(app: FastAPI) -> AsyncGenerator[None, None]:
        if model is not None:
            ml_models["credit_default"] = model
        else:
            from . import load_model

            ml_models["credit_default"] = (loader or load_model)(os.getenv("MODEL_DIRECTORY", "./app/model"))
        m = ml_models["credit_default"]
        ml_models["_batcher"] = MicroBatcher(
            getattr(m, "replicas", None) or [m],
            max_rows=int(os.environ.get("B200_MAX_BATCH", "65536")),
            window_us=int(os.environ.get("B200_BATCH_WINDOW_US", "200")),
        )
        yield
        ml_models["_batcher"].close()
        closer = getattr(ml_models.get("credit_default"), "close", None)
        ml_models.clear()
        if closer and model is None:
            closer()

    app = FastAPI(title=_service_name(), docs_url="/", lifespan=lifespan)

    @app.post("/predict", response_model=ModelOutput, openapi_extra={"requestBody": _REQUEST_SCHEMA})
    async def predict(request: Request):
        """Score a list of loan applicants: default probability, outlier flag, per-feature batch drift."""
        # list[LoanApplicant] semantics (422 on a type error, defaults filled): native one-pass parser for bodies of the
        # regular shape, pydantic-core for everything else (ingest.py)
        input_df = parser.frame(await request.body())
        if len(input_df) == 0:
            # the reference's empty DataFrame has no columns and dies in df[self.all_features] -> HTTP 500
            raise KeyError(f"None of {ALL_FEATURES} are in the [columns]")
        m = ml_models["credit_default"]
        request_id = uuid.uuid4().hex
        log_pool.submit(lambda: _log_record("InferenceData", request_id, input_df.to_json(orient="records")))

        # the drift scores depend on this request's rows only (no cross-request batching): start them first, on their
        # own device stream, and let them run while the batcher encodes and scores the rows
        drift = getattr(m, "drift", None)
        pending = asyncio.get_running_loop().run_in_executor(None, drift.score, input_df) if drift is not None else None
        try:
            proba, flags = await ml_models["_batcher"].score(input_df)
        finally:
            drift_scores = (await pending) if pending is not None else [0.0] * len(ALL_FEATURES)
        model_output = {
            "predictions": _as_list(proba),
            "outliers": _as_list(flags) if flags is not None else [0] * len(input_df),
            "feature_drift_batch": dict(zip(ALL_FEATURES, drift_scores)),
        }
        log_pool.submit(_log_record, "ModelOutput", request_id, model_output)
        # Response side (reference app/main.py:42,86: `response_model=ModelOutput` makes FastAPI validate the dict field by
        # field and run it through jsonable_encoder before rendering).  Every value here was produced by this handler with
        # the declared types, so the body is rendered once, exactly as Starlette's JSONResponse would render the validated
        # model (compact separators, allow_nan=False -> a NaN drift score is the same ValueError -> HTTP 500)
        body = json.dumps(model_output, ensure_ascii=False, allow_nan=False, indent=None, separators=(",", ":")).encode("utf-8")
        return Response(content=body, media_type="application/json")

    return app


logging.basicConfig(level=logging.INFO)

if __name__ == "__main__":
    import uvicorn

    uvicorn.run(create_app(), host="0.0.0.0", port=5000)
