"""Constants of the node layer, under the names the reference exports from ``cbn.base`` (``cbn/base/__init__.py``).

``KEY_MAX_CARDINALITY_FOR_DISCRETE`` is the keyword a caller passes to ``BayesianNetwork`` / ``Node`` to say how many
distinct values still make a column "discrete"; ``BASE_MAX_CARDINALITY`` is its default.  The engine itself supports up
to 255 distinct values per variable (uint8 codes); the two ``KEY_*`` kind labels are what ``Node.info`` records.
"""
BASE_MAX_CARDINALITY = 20
KEY_MAX_CARDINALITY_FOR_DISCRETE = "max_cardinality_for_discrete_domain"
KEY_DISCRETE, KEY_CONTINUOUS = "discrete", "continuous"

__all__ = ["BASE_MAX_CARDINALITY", "KEY_MAX_CARDINALITY_FOR_DISCRETE", "KEY_DISCRETE", "KEY_CONTINUOUS"]
