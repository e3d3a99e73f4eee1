"""Result types of the KLLMs client surface."""
from ._models import KLLMsChatCompletion, KLLMsParsedChatCompletion

__all__ = ["KLLMsChatCompletion", "KLLMsParsedChatCompletion"]
