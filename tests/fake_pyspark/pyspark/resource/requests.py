"""Fake pyspark.resource.requests."""


class TaskResourceRequests:
    def __init__(self):
        self.req = {}

    def cpus(self, n):
        self.req["cpus"] = n
        return self

    def resource(self, name, amount):
        self.req[name] = amount
        return self
