from .miller_schupp import (generate_miller_schupp_presentations,  # noqa: F401
                            trivialize_miller_schupp_through_search, write_list_to_text_file)
