"""Stub of the third-party ``termcolor`` package (absent from this image).

Test infrastructure only: lets the unmodified reference (`/root/reference/demo1/util.py:4`)
import in this container.  See SURVEY.md F3.
"""


def colored(text, *args, **kwargs):
    return text
