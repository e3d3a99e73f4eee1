from .completions import KLLMsChatCompletion
from .parsed import KLLMsParsedChatCompletion

__all__ = ["KLLMsParsedChatCompletion", "KLLMsChatCompletion"]
