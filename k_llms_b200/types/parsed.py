"""Import location kept for callers of the reference layout (k_llms.types.parsed)."""
from ._models import KLLMsParsedChatCompletion  # noqa: F401
