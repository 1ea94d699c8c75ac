"""Names imported by fewshot_lsr.py at module level; unused by the functions the fixture runs."""


class NamedSharding:
  pass


class PartitionSpec:
  pass


class Mesh:
  pass
