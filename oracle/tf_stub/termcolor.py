"""Stand-in for `termcolor` (imported by the reference's utils.py for log_args only); test infrastructure."""


def colored(text, *args, **kwargs):
    return text
