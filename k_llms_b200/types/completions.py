"""Import location kept for callers of the reference layout (k_llms.types.completions)."""
from ._models import KLLMsChatCompletion  # noqa: F401
