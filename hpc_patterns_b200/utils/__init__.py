"""Timing, clocks sampling, log parsing, reports and the dtype table."""
