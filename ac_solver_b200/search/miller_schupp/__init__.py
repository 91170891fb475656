from .miller_schupp import (generate_miller_schupp_presentations,  # noqa: F401
                            load_initial_states_from_text_file, load_presentations_from_text_file,
                            trivialize_miller_schupp_through_search, write_list_to_text_file)
