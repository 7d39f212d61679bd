"""CPU: the restated configs (indm_b200/configs.py) equal the reference's config objects leaf by leaf."""
import pytest

from helpers import load_json
from indm_b200 import configs


@pytest.mark.parametrize("name", configs.available())
def test_config_matches_reference(name):
    want = load_json('configs.json')[name]
    got = configs.get_config(name).to_plain()
    got.pop('device')
    got['flow'].pop('wolf_params')   # restated JSON content, checked in test_flow_* against the reference JSON
    assert got == want
