"""ParsedChatCompletion + `likelihoods` (reference k_llms/types/parsed.py:8-15)."""
from typing import Any, Dict, Optional

from openai.types.chat import ParsedChatCompletion
from pydantic import Field


class KLLMsParsedChatCompletion(ParsedChatCompletion):
    likelihoods: Optional[Dict[str, Any]] = Field(
        default=None,
        description="Per-field confidence of the consensus, same structure as the extraction object.",
    )
