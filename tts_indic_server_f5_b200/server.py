"""`POST /v1/audio/speech` of the Dhwani server on the batching engine (the HTTP side of the path).

Mirrors `src/server/routes/speech.py:19-41` (body `{"text": ...}`; 503 while `tts_manager.model` is falsy, 400 for empty text,
an `audio/wav` attachment named `synthesized_kannada_speech.wav`), `src/server/utils/tts_utils.py:39-65` (`synthesize_speech`:
voice looked up by name, reference text defaulting to the voice's transcript, the three 400s, int16 -> float rescale, 16-bit PCM
WAV at 24 kHz), the `X-Response-Time` middleware of `src/server/main.py:77-86` and `GET /v1/health` (`routes/health.py:9-11`).

Two deliberate differences, both on the serving side of the boundary:
  * the prompt is a LOCAL wav registered at start-up (`Voice`); the reference downloads it from GitHub on every request
    (`tts_utils.py:31-36,54`) and re-writes it to a temp file that is never unlinked (`:55-58`);
  * the handler does not run the model on the event-loop thread (`speech.py:19` is `async def` with a blocking body, so the
    reference serves one request at a time): it awaits the `ContinuousScheduler`'s future, and requests that arrive while
    the GPU is busy are sampled together as one packed batch.  `batching=False` keeps the manager call
    (`tts_manager.synthesize`, boundary #1) on a worker thread instead.

FastAPI is only needed by this module; nothing else in the package imports it."""
from __future__ import annotations

import asyncio
import queue
import time
from contextlib import asynccontextmanager
from dataclasses import dataclass

import numpy as np
from fastapi import FastAPI, HTTPException, Request
from pydantic import BaseModel
from starlette.concurrency import run_in_threadpool
from starlette.responses import StreamingResponse


@dataclass
class Voice:
    """One row of the reference's EXAMPLES table (`tts_utils.py:12-19`): prompt audio (a local PCM wav) and its transcript."""
    audio_path: str
    ref_text: str


class KannadaSynthesizeRequest(BaseModel):      # tts_utils.py:27-28
    text: str


class SynthesizeRequest(BaseModel):             # tts_utils.py:22-25
    text: str
    ref_audio_name: str
    ref_text: str | None = None


def create_app(tts_manager, voices: dict[str, Voice], default_voice: str = "KAN_F (Happy)", batching: bool = True,
               max_queue: int = 256, max_batch_requests: int = 64, max_wait_ms: float = 4.0) -> FastAPI:
    """`tts_manager`: `api.TTSManager` (or anything with `.model`, `.load()`, `.synthesize(text, ref_audio_path, ref_text)`).
    The model is loaded in the lifespan handler like `main.py:37-57`; the scheduler is built on first use and closed on
    shutdown."""
    if default_voice not in voices:
        raise ValueError(f"default voice {default_voice!r} is not in the voice table")
    state: dict = {"sched": None}

    @asynccontextmanager
    async def lifespan(app: FastAPI):
        if not tts_manager.model:
            await run_in_threadpool(tts_manager.load)
        yield
        if state["sched"] is not None:
            state["sched"].close()
            state["sched"] = None

    app = FastAPI(title="Dhwani API (B200-native F5-TTS path)", version="1.0.0", redirect_slashes=False, lifespan=lifespan)

    @app.middleware("http")
    async def add_request_timing(request: Request, call_next):
        t0 = time.time()
        response = await call_next(request)
        response.headers["X-Response-Time"] = f"{time.time() - t0:.3f}"
        return response

    def scheduler():
        if state["sched"] is None:
            from .api import _synthesizer_for
            from .scheduler import ContinuousScheduler
            m = tts_manager.model
            state["sched"] = ContinuousScheduler(_synthesizer_for(m.ema_model, m.vocoder), max_queue=max_queue,
                                                 max_batch_requests=max_batch_requests, max_wait_ms=max_wait_ms)
        return state["sched"]

    async def synthesize_speech(text: str, ref_audio_name: str, ref_text: str | None):
        from .api import wav_response_bytes
        voice = voices.get(ref_audio_name)
        if voice is None:
            raise HTTPException(status_code=400, detail="Invalid reference audio name.")
        if not ref_text:
            ref_text = voice.ref_text
        if not text.strip():
            raise HTTPException(status_code=400, detail="Text to synthesize cannot be empty.")
        if not ref_text or not ref_text.strip():
            raise HTTPException(status_code=400, detail="Reference text cannot be empty.")
        model = tts_manager.model
        if batching and hasattr(model, "_prompt"):
            ref_audio, cond_text = await run_in_threadpool(model._prompt, voice.audio_path, ref_text)   # cached per voice
            try:
                fut = scheduler().submit(ref_audio, cond_text, text, timeout=0)
            except queue.Full:
                raise HTTPException(status_code=503, detail="TTS queue is full") from None
            try:
                wave, _, _ = await asyncio.wrap_future(fut)
            except RuntimeError as e:
                if hasattr(tts_manager, "note_failure"):
                    tts_manager.note_failure(e)                       # a CUDA error is sticky: the readiness probe turns 503
                raise
            audio = np.clip(wave * 32768.0, -32768, 32767).astype(np.int16) if getattr(model, "output_int16", True) else wave
        else:
            audio = await run_in_threadpool(tts_manager.synthesize, text, voice.audio_path, ref_text)
        return wav_response_bytes(np.asarray(audio))

    def wav_attachment(buf, name: str) -> StreamingResponse:
        return StreamingResponse(buf, media_type="audio/wav", headers={"Content-Disposition": f"attachment; filename={name}"})

    @app.post("/v1/audio/speech", response_class=StreamingResponse)
    async def synthesize_kannada(request: KannadaSynthesizeRequest):
        if not tts_manager.model:
            raise HTTPException(status_code=503, detail="TTS model not loaded")
        if not request.text.strip():
            raise HTTPException(status_code=400, detail="Text to synthesize cannot be empty.")
        buf = await synthesize_speech(request.text, default_voice, voices[default_voice].ref_text)
        return wav_attachment(buf, "synthesized_kannada_speech.wav")

    @app.post("/v1/audio/speech/voice", response_class=StreamingResponse)
    async def synthesize_voice(request: SynthesizeRequest):
        """The general form `synthesize_speech` implements but no reference route exposes (`SynthesizeRequest`, tts_utils.py:22-25)."""
        if not tts_manager.model:
            raise HTTPException(status_code=503, detail="TTS model not loaded")
        buf = await synthesize_speech(request.text, request.ref_audio_name, request.ref_text)
        return wav_attachment(buf, "synthesized_speech.wav")

    @app.get("/v1/health")
    async def health_check():
        sched = state["sched"]
        return {"status": "healthy" if tts_manager.model else "unavailable", "model": getattr(tts_manager, "repo_id", "ai4bharat/IndicF5"),
                "failed_reason": getattr(tts_manager, "failed_reason", None),
                "batches": list(sched.batches[-16:]) if sched is not None else []}

    return app
