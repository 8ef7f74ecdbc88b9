from xml.etree.ElementTree import *  # noqa: F401,F403
