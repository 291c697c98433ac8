"""Stand-in for lxml: the reference only uses lxml.etree.parse(...).getroot() and element find/findall/get/attrib."""
