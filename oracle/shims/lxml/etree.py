import gzip
import xml.etree.ElementTree as _ET

Element = _ET.Element


class _Comment:  # stdlib ElementTree drops comments while parsing, so nothing is ever an instance
    pass


def parse(path):
    path = str(path)
    if path.endswith(".gz"):
        with gzip.open(path, "rb") as f:
            return _ET.parse(f)
    return _ET.parse(path)


fromstring = _ET.fromstring
tostring = _ET.tostring
