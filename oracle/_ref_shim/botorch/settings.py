class debug:
    @classmethod
    def on(cls):
        return False
