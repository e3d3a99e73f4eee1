"""ChatCompletion + `likelihoods` (reference k_llms/types/completions.py:8-15)."""
from typing import Any, Dict, Optional

from openai.types.chat import ChatCompletion
from pydantic import Field


class KLLMsChatCompletion(ChatCompletion):
    likelihoods: Optional[Dict[str, Any]] = Field(
        default=None,
        description="Per-field confidence of the consensus, same structure as the extraction object.",
    )
