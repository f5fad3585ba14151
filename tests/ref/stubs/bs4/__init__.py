# Stub so that /root/reference/GRiD/URDFParser/URDFParser.py imports (it only uses xml.etree at run time).
BeautifulSoup = None
