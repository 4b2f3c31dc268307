class HTTPClient:
    pass
