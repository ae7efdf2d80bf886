"""Exception types of the dspeed API (reference: src/dspeed/errors.py:1-40)."""

from __future__ import annotations


class DSPError(Exception):
    """Base class for signal-processing errors."""


class DSPFatal(DSPError):
    """Fatal error raised by a processor; halts production.

    ``wf_range`` (entries being processed) and ``processor`` (processor and arguments)
    are filled in by the processing chain after the exception is caught and appended to
    the message, like the reference does (processing_chain.py:1156-1159)."""

    def __init__(self, *args) -> None:
        super().__init__(*args)
        self.wf_range = None
        self.processor = None
        self.code = None

    def __str__(self) -> str:
        suffix = ""
        if self.wf_range:
            suffix += "\nThrown while processing entries " + str(self.wf_range)
        if self.processor:
            suffix += "\nThrown by " + self.processor
        return super().__str__() + suffix


class ProcessingChainError(DSPError):
    """Problem while setting up a processing chain."""
