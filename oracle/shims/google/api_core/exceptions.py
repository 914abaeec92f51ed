class ServerError(Exception):
    pass


class Forbidden(Exception):
    pass


class ServiceUnavailable(Exception):
    pass
