"""`from lxml import etree` -> the standard library's ElementTree (the reference only parses URDF strings with it)."""
from xml.etree import ElementTree as etree  # noqa: F401
