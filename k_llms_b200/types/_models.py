"""The two result types of the client surface: OpenAI's completion models plus one `likelihoods` field
(reference k_llms/types/completions.py:8-15 and k_llms/types/parsed.py:8-15 declare the same field on each)."""
from typing import Any, Dict, Optional

from openai.types.chat import ChatCompletion, ParsedChatCompletion
from pydantic import Field, create_model

_LIKELIHOODS = (Optional[Dict[str, Any]],
                Field(default=None, description="Per-field confidence of the consensus, same structure as the extraction object."))


def _with_likelihoods(name: str, base):
    model = create_model(name, __base__=base, __module__=__name__, likelihoods=_LIKELIHOODS)
    model.__doc__ = f"{base.__name__} + `likelihoods`: per-field confidences of the consensus in choices[0]."
    return model


KLLMsChatCompletion = _with_likelihoods("KLLMsChatCompletion", ChatCompletion)
KLLMsParsedChatCompletion = _with_likelihoods("KLLMsParsedChatCompletion", ParsedChatCompletion)
