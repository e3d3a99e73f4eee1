"""`KLLMs` / `AsyncKLLMs`: OpenAI client wrappers exposing `.chat.completions.create()/.parse()` with n-way
consensus (reference k_llms/client.py:15-196).  Host plumbing only — the network calls are unchanged; what changes
is where the n choices are consolidated (the GPU, see utils/consensus_utils.py)."""
from __future__ import annotations

import asyncio
import os
from typing import Any, Awaitable, Callable, List, Optional

from openai import AsyncOpenAI, OpenAI

from .resources.completions import AsyncCompletions, Completions

MAX_TOKENS_PER_MODEL = {"text-embedding-3-small": 8191, "text-embedding-3-large": 8191}
PRICING = {"text-embedding-3-small": 0.020, "text-embedding-3-large": 0.13}  # USD per 1M tokens


class BaseOpenAIWrapper:
    def __init__(self, api_key: Optional[str] = None, base_url: Optional[str] = None, timeout: Optional[float] = None,
                 max_retries: int = 2, **kwargs: Any):
        self.api_key = api_key or os.environ.get("OPENAI_API_KEY")
        self.base_url = base_url
        self.timeout = timeout
        self.max_retries = max_retries
        self._extra_kwargs = kwargs

    def _client_kwargs(self) -> dict:
        return dict(api_key=self.api_key, base_url=self.base_url, timeout=self.timeout, max_retries=self.max_retries,
                    **self._extra_kwargs)


class Chat:
    def __init__(self, wrapper: "KLLMs"):
        self._wrapper = wrapper
        self.completions = Completions(wrapper)


class AsyncChat:
    def __init__(self, wrapper: "AsyncKLLMs"):
        self._wrapper = wrapper
        self.completions = AsyncCompletions(wrapper)


class KLLMs(BaseOpenAIWrapper):
    def __init__(self, **kwargs: Any):
        super().__init__(**kwargs)
        self._client = OpenAI(**self._client_kwargs())
        self.chat = Chat(self)
        self.get_embeddings: Callable[[List[str], str, int, bool], List[List[float]]] = (
            lambda texts, model, batch_size, verbose: get_embeddings(self._client, texts, model, batch_size, verbose))

    @property
    def client(self) -> OpenAI:
        return self._client


class AsyncKLLMs(BaseOpenAIWrapper):
    def __init__(self, **kwargs: Any):
        super().__init__(**kwargs)
        self._client = AsyncOpenAI(**self._client_kwargs())
        self.chat = AsyncChat(self)
        self.get_embeddings: Callable[[List[str], str, int, bool], Awaitable[List[List[float]]]] = (
            lambda texts, model, batch_size, verbose: async_get_embeddings(self._client, texts, model, batch_size, verbose))

    @property
    def client(self) -> AsyncOpenAI:
        return self._client


def _crop(texts: List[str], model: str, only_long: bool) -> List[str]:
    """Trim each text to the model's token window (tiktoken), optionally only texts long enough to matter."""
    import tiktoken
    enc = tiktoken.encoding_for_model(model)
    limit = MAX_TOKENS_PER_MODEL[model]
    return [enc.decode(enc.encode(t)[:limit]) if (not only_long or len(t) * 3 > limit) else t for t in texts]


def _check_model(model: str) -> None:
    if model not in MAX_TOKENS_PER_MODEL:
        raise ValueError(f"Model {model} not supported. Available models: {list(MAX_TOKENS_PER_MODEL.keys())}")


def get_embeddings(openai_client: OpenAI, texts: List[str], model: str = "text-embedding-3-small", batch_size: int = 2048,
                   verbose: bool = False) -> List[List[float]]:
    """Batched OpenAI embeddings (reference client.py:75-122)."""
    _check_model(model)
    texts = _crop(texts, model, only_long=False)
    out: List[List[float]] = []
    spent = 0.0
    for start in range(0, len(texts), batch_size):
        response = openai_client.embeddings.create(input=texts[start:start + batch_size], model=model)
        spent += response.usage.prompt_tokens * PRICING[model] / 1_000_000.0
        out.extend(item.embedding for item in response.data)
    if verbose:
        print(f"TOTAL PRICE: ${spent:.6f}")
    return out


async def async_get_embeddings(openai_client: AsyncOpenAI, texts: List[str], model: str = "text-embedding-3-small",
                               batch_size: int = 2048, verbose: bool = False) -> List[List[float]]:
    """Async twin with the reference's retry-with-everything-cropped fallback (client.py:125-196)."""
    _check_model(model)

    async def embed_all(batch_source: List[str]) -> List[List[float]]:
        out: List[List[float]] = []
        for start in range(0, len(batch_source), batch_size):
            response = await openai_client.embeddings.create(input=batch_source[start:start + batch_size], model=model)
            out.extend(item.embedding for item in response.data)
        return out

    try:
        return await embed_all(await asyncio.to_thread(_crop, texts, model, True))
    except Exception as exc:
        if verbose:
            print(f"Embedding request failed with error: {exc}. Retrying with all strings cropped.")
        return await embed_all(await asyncio.to_thread(_crop, texts, model, False))
